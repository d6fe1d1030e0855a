"""Mirror of reference models/anovaegan.py (same name, signature anovaegan(x, dropout_rate, dropout, config) and output keys)."""
from .customlayers import GraphSpec, GraphTensor, build_unified_decoder, build_unified_encoder

KEYS = ('z_mu', 'z_log_sigma', 'z_sigma', 'out', 'd_fake_features', 'd_', 'd_features', 'd', 'x_hat', 'd_hat_features', 'd_hat')


def anovaegan(x, dropout_rate, dropout, config):
    shape = x.get_shape().as_list()
    encoder = build_unified_encoder(shape, config.intermediateResolutions)
    generator = build_unified_decoder(config.outputWidth, config.intermediateResolutions, config.numChannels, use_batchnorm=False)
    graph = GraphSpec('anovaegan', shape, config, encoder, generator)
    graph.discriminator = build_unified_encoder(shape, config.intermediateResolutions, use_batchnorm=False)
    return {key: GraphTensor(graph, key) for key in KEYS}
