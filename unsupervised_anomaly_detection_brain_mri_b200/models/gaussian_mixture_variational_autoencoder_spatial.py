"""Mirror of reference models/gaussian_mixture_variational_autoencoder_spatial.py (same name, signature and output keys)."""
import types

from .customlayers import GraphSpec, GraphTensor, build_unified_decoder, build_unified_encoder
from .gaussian_mixture_variational_autoencoder import KEYS


def gaussian_mixture_variational_autoencoder_spatial(x, dropout_rate, dropout, config):
    shape = x.get_shape().as_list()
    encoder = build_unified_encoder(shape, config.intermediateResolutions)
    decoder = build_unified_decoder(config.outputWidth, config.intermediateResolutions, config.numChannels)
    latent = types.SimpleNamespace(zDim=config.dim_z, intermediateResolutions=config.intermediateResolutions)
    graph = GraphSpec('gaussian_mixture_variational_autoencoder_spatial', shape, latent, encoder, decoder)
    graph.dim_w, graph.dim_c = int(config.dim_w), int(config.dim_c)
    return {key: GraphTensor(graph, key) for key in KEYS}
