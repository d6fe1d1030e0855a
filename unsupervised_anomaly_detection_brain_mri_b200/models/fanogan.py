"""Mirror of reference models/fanogan.py (same name, signature fanogan(z, x, dropout_rate, dropout, config) and output keys)."""
from .customlayers import GraphSpec, GraphTensor, build_unified_decoder, build_unified_encoder

KEYS = ('z_enc', 'x_enc', 'x_', 'd_fake_features', 'd_', 'd_features', 'd', 'x_hat', 'd_hat_features', 'd_hat',
        'd_enc_features', 'd_enc')


def fanogan(z, x, dropout_rate, dropout, config):
    shape = x.get_shape().as_list()
    encoder = build_unified_encoder(shape, config.intermediateResolutions)
    generator = build_unified_decoder(config.outputWidth, config.intermediateResolutions, config.numChannels, use_batchnorm=False)
    graph = GraphSpec('fanogan', shape, config, encoder, generator)
    graph.discriminator = build_unified_encoder(shape, config.intermediateResolutions, use_batchnorm=False)
    return {key: GraphTensor(graph, key) for key in KEYS}
