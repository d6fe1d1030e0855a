#!/usr/bin/env python
"""Operand traffic (L2 -> shared memory) that the tcgen05 gather kernels move per launch, derived from the layer shapes, beside the
measured kernel times of a bench layer table.  A k-block = one filter tap x 32 input channels: one 16 KB
activation tile (128 pixels x 128 B) + one {hi, lo} weight image of 2 * N * 128 B.  Tensor work per k-block from
profiles/r1_ubench_mma_rate.txt (A from tensor memory, paired scheme at N <= 64): 206 / 384 / 768 cycles at N = 32 / 64 / 128.
usage: python tools/l2_traffic_model.py [layer-table.json] [--cap-bytes-per-clk 6300] [--mhz 1965]  ->  JSON on stdout"""
import argparse
import json
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TENSOR_CYCLES = {32: 206, 64: 384, 128: 768}


def vae_layers(S=256, B=64, res=8):
    """(name, op) -> (pixels M of the GEMM, K-blocks per tile, N) for the VAE stack (reference models/customlayers.py:16-38)."""
    n = int(math.log2(S) - math.log2(res))
    enc = [min(128, 32 * 2 ** i) for i in range(n)]
    dec = [max(32, 128 // 2 ** i) for i in range(n)]
    out = {}
    cin, s = 1, S
    for i, co in enumerate(enc):
        so = s // 2
        if cin >= 32:
            out[f'enc_conv2D_{i}:conv2d_fwd'] = (B * so * so, 25 * cin // 32, co)              # Form F over x
            out[f'enc_conv2D_{i}:conv2d_dgrad'] = (B * so * so * 4, 25 * co // 32 / 4, cin)    # Form T over dz: 4 classes, 25 taps in all
        cin, s = co, so
    for i, co in enumerate(dec):
        out[f'dec_Conv2DT_{i}:convT2d_fwd'] = (B * s * s * 4, 25 * cin // 32 / 4, co)          # Form T over x
        out[f'dec_Conv2DT_{i}:convT2d_dgrad'] = (B * s * s, 25 * co // 32, cin)                # Form F over dy
        cin, s = co, s * 2
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('table', nargs='?', default=os.path.join(ROOT, 'profiles', 'r1_layers_tc3_v16.json'))
    ap.add_argument('--cap-bytes-per-clk', type=float, default=6300.0, help='chip-wide L2 slice throughput cap (B300 notes; measure with l2_to_sm)')
    ap.add_argument('--mhz', type=float, default=1965.0)
    ap.add_argument('--sms', type=int, default=148)
    a = ap.parse_args()
    ms = {r['op']: r['ms'] for r in json.load(open(a.table))['table']}
    cap_tbs = a.cap_bytes_per_clk * a.mhz * 1e6 / 1e12
    rows = []
    for op, (M, kb_per_tile, N) in vae_layers().items():
        if op not in ms or N not in TENSOR_CYCLES:
            continue
        kblocks = M / 128 * kb_per_tile
        per_kb = 16384 + 2 * N * 128
        gb = kblocks * per_kb / 1e9
        t = ms[op] * 1e-3
        cyc_per_kb = t * a.mhz * 1e6 * a.sms / kblocks
        rows.append(dict(op=op, N=N, kblocks=int(kblocks), kb_bytes=per_kb, l2_to_sm_gb=round(gb, 3), ms=round(ms[op], 4),
                         achieved_tbs=round(gb / 1e3 / t, 2), frac_of_cap=round(gb / 1e3 / t / cap_tbs, 3),
                         ms_at_cap=round(gb / 1e3 / cap_tbs * 1e3, 4), cycles_per_kblock_per_sm=round(cyc_per_kb),
                         cycles_at_cap=round(per_kb / (a.cap_bytes_per_clk / a.sms)), tensor_cycles=TENSOR_CYCLES[N],
                         max_tensor_busy_at_cap=round(TENSOR_CYCLES[N] / (per_kb / (a.cap_bytes_per_clk / a.sms)), 3)))
    rows.sort(key=lambda r: -r['ms'])
    print(json.dumps(dict(source=os.path.basename(a.table), cap_bytes_per_clk=a.cap_bytes_per_clk, cap_tbs=round(cap_tbs, 2), mhz=a.mhz,
                          note='traffic is derived from shapes (every k-block loads its own tiles), times are measured', rows=rows), indent=1))


if __name__ == '__main__':
    main()
