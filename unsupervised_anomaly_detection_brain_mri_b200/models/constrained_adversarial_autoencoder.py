"""Mirror of reference models/constrained_adversarial_autoencoder.py (same name, signature and output keys)."""
from .customlayers import GraphSpec, GraphTensor, build_unified_decoder, build_unified_encoder

KEYS = ('z_', 'x_hat', 'z_rec', 'd_', 'd', 'z_hat', 'd_hat')


def constrained_adversarial_autoencoder(z, x, dropout_rate, dropout, config):
    shape = x.get_shape().as_list()
    encoder = build_unified_encoder(shape, config.intermediateResolutions)
    decoder = build_unified_decoder(config.outputWidth, config.intermediateResolutions, config.numChannels)
    graph = GraphSpec('constrained_adversarial_autoencoder', shape, config, encoder, decoder)
    graph.discriminator = [{'op': 'dense', 'units': 100, 'activation': 'leaky_relu', 'alpha': 0.2},
                           {'op': 'dense', 'units': 50, 'activation': 'leaky_relu', 'alpha': 0.2}, {'op': 'dense', 'units': 1}]
    return {key: GraphTensor(graph, key) for key in KEYS}
