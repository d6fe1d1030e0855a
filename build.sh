#!/bin/bash
# Build libuad_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
PKG=unsupervised_anomaly_detection_brain_mri_b200
SRC="$PKG/csrc/uad_conv_simt.cu $PKG/csrc/uad_conv_api.cu $PKG/csrc/uad_conv_tc.cu $PKG/csrc/uad_conv_hs.cu $PKG/csrc/uad_conv_ws.cu $PKG/csrc/uad_dense.cu $PKG/csrc/uad_elementwise.cu $PKG/csrc/uad_scoring.cu $PKG/csrc/uad_fanogan.cu $PKG/csrc/uad_restore.cu $PKG/csrc/uad_gmvae.cu $PKG/csrc/uad_peer.cu"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $SRC -o $PKG/libuad_b200.so -lcuda "$@"
echo "built $PKG/libuad_b200.so"
