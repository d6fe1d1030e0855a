"""Mirror of reference models/gaussian_mixture_variational_autoencoder.py (same name, signature and output keys)."""
import types

from .customlayers import GraphSpec, GraphTensor, build_unified_decoder, build_unified_encoder

KEYS = ('w_mu', 'w_log_sigma', 'w_sampled', 'z_mu', 'z_log_sigma', 'z_sampled', 'z_wc_mus', 'z_wc_log_sigma_invs', 'z_wc_sampled', 'xz_mu',
        'pc_logit', 'pc')


def gaussian_mixture_variational_autoencoder(x, dropout_rate, dropout, config):
    shape = x.get_shape().as_list()
    encoder = build_unified_encoder(shape, config.intermediateResolutions)
    decoder = build_unified_decoder(config.outputWidth, config.intermediateResolutions, config.numChannels)
    latent = types.SimpleNamespace(zDim=config.dim_z, intermediateResolutions=config.intermediateResolutions)   # z has dim_z entries here
    graph = GraphSpec('gaussian_mixture_variational_autoencoder', shape, latent, encoder, decoder)
    graph.dim_w, graph.dim_c = int(config.dim_w), int(config.dim_c)
    return {key: GraphTensor(graph, key) for key in KEYS}
