"""Mirror of reference models/constrained_autoencoder.py (same name, signature and output keys): the dense AE whose
reconstruction is mapped back to the latent space by the SAME encoder layers (z_rec), both Dropout calls of the
bottleneck honouring the flag (:29-30) - unlike models/autoencoder.py."""
from .customlayers import GraphSpec, GraphTensor, build_unified_decoder, build_unified_encoder


def constrained_autoencoder(x, dropout_rate, dropout, config):
    shape = x.get_shape().as_list()
    encoder = build_unified_encoder(shape, config.intermediateResolutions)
    decoder = build_unified_decoder(config.outputWidth, config.intermediateResolutions, config.numChannels)
    graph = GraphSpec('constrained_autoencoder', shape, config, encoder, decoder)
    return {key: GraphTensor(graph, key) for key in 'z,x_hat,z_rec'.split(',')}
