"""Shared encoder / decoder stacks (mirror of reference models/customlayers.py:16-38) as layer descriptors."""
import math


class Placeholder:
    """Stand-in for tf.placeholder: carries the static NHWC shape ([None, H, W, C])."""

    def __init__(self, shape, name='x'):
        self.shape = list(shape)
        self.name = name

    def get_shape(self):
        return self

    def as_list(self):
        return list(self.shape)


class GraphTensor:
    """Handle of one named output of a GraphSpec (what the reference returns as a tf.Tensor)."""

    def __init__(self, graph, key):
        self.graph = graph
        self.key = key

    def __repr__(self):
        return f'<GraphTensor {self.graph.arch}:{self.key}>'


class GraphSpec:
    def __init__(self, arch, input_shape, config, encoder, decoder):
        self.arch = arch
        self.S = int(input_shape[1])
        self.C = int(input_shape[3])
        self.zDim = int(config.zDim)
        self.res = int(config.intermediateResolutions[0])
        self.encoder = encoder
        self.decoder = decoder


def build_unified_encoder(input_shape, intermediateResolutions, use_batchnorm=True):
    """n x [Conv2D(min(128, 32*2^i), k=5, s=2, 'same') -> BatchNormalization|LayerNormalization([1,2]) -> LeakyReLU()]."""
    encoder = []
    num_pooling = int(math.log(input_shape[1], 2) - math.log(float(intermediateResolutions[0]), 2))
    for i in range(num_pooling):
        filters = int(min(128, 32 * (2 ** i)))
        encoder.append({'op': 'conv2d', 'filters': filters, 'kernel_size': 5, 'strides': 2, 'name': f'enc_conv2D_{i}'})
        encoder.append({'op': 'batchnorm' if use_batchnorm else 'layernorm_hw'})
        encoder.append({'op': 'leaky_relu', 'alpha': 0.3})
    return encoder


def build_unified_decoder(outputWidth, intermediateResolutions, outputChannels, final_activation=None, use_batchnorm=True):
    """BN -> ReLU -> n x [Conv2DTranspose(max(32, 128/2^i), 5, s=2, 'same') -> BN -> LeakyReLU()] -> Conv2D(C, 1)."""
    decoder = []
    num_upsampling = int(math.log(outputWidth, 2) - math.log(float(intermediateResolutions[0]), 2))
    decoder.append({'op': 'batchnorm' if use_batchnorm else 'layernorm_hw'})
    decoder.append({'op': 'relu'})
    for i in range(num_upsampling):
        filters = int(max(32, 128 / (2 ** i)))
        decoder.append({'op': 'conv2d_transpose', 'filters': filters, 'kernel_size': 5, 'strides': 2, 'name': f'dec_Conv2DT_{i}'})
        decoder.append({'op': 'batchnorm' if use_batchnorm else 'layernorm_hw'})
        decoder.append({'op': 'leaky_relu', 'alpha': 0.3})
    decoder.append({'op': 'conv2d', 'filters': outputChannels, 'kernel_size': 1, 'strides': 1, 'name': 'dec_Conv2D_final',
                    'activation': final_activation})
    return decoder
