"""Mirror of reference models/autoencoder_spatial.py (same name, signature and output keys): the unified encoder, Dropout on
the spatial code z [B, res, res, C], the unified decoder - no dense bottleneck."""
from .customlayers import GraphSpec, GraphTensor, build_unified_decoder, build_unified_encoder


def autoencoder_spatial(x, dropout_rate, dropout, config):
    shape = x.get_shape().as_list()
    encoder = build_unified_encoder(shape, config.intermediateResolutions)
    decoder = build_unified_decoder(config.outputWidth, config.intermediateResolutions, config.numChannels)
    graph = GraphSpec('autoencoder_spatial', shape, config, encoder, decoder)
    return {key: GraphTensor(graph, key) for key in 'z,x_hat'.split(',')}
