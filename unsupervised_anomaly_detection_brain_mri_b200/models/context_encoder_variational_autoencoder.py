"""Mirror of reference models/context_encoder_variational_autoencoder.py (same name, signature and output keys)."""
from .customlayers import GraphSpec, GraphTensor, build_unified_decoder, build_unified_encoder


def context_encoder_variational_autoencoder(x, x_ce, dropout_rate, dropout, config):
    shape = x.get_shape().as_list()
    encoder = build_unified_encoder(shape, config.intermediateResolutions)
    decoder = build_unified_decoder(config.outputWidth, config.intermediateResolutions, config.numChannels)
    graph = GraphSpec('context_encoder_variational_autoencoder', shape, config, encoder, decoder)
    return {key: GraphTensor(graph, key) for key in 'z_mu,z_mu_ce,z_log_sigma,z_sigma,x_hat,x_hat_ce'.split(',')}
