"""Developer aid: filter-gradient error of the tensor-core Form-W kernels against float64 at VAE-256 layer shapes, as a function of
the accumulator depth (UAD_WS_DEPTH, read once per process: run one depth per process)."""
import os, sys
import torch
sys.path.insert(0, '.')
from unsupervised_anomaly_detection_brain_mri_b200 import abi
from unsupervised_anomaly_detection_brain_mri_b200.abi import call
L = abi.lib()
DEV = 'cuda:0'
st = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
cases = [('enc2 conv', 'conv', 64, 64, 64, 128), ('enc1 conv', 'conv', 64, 128, 32, 64), ('dec4 convT', 'convT', 64, 128, 32, 32), ('enc3 conv', 'conv', 64, 32, 128, 128)]
for name, kind, B, H, Cin, Cout in cases:
    if kind == 'conv':
        x = torch.rand(B, H, H, Cin, device=DEV); dz = torch.randn(B, H // 2, H // 2, Cout, device=DEV) * 0.1
        dw = torch.empty(5, 5, Cin, Cout, device=DEV); op = 2
    else:
        x = torch.rand(B, H, H, Cin, device=DEV); dz = torch.randn(B, 2 * H, 2 * H, Cout, device=DEV) * 0.1
        dw = torch.empty(5, 5, Cout, Cin, device=DEV); op = 5
    wsb = L.uad_conv_workspace_bytes(op, B, H, H, Cin, Cout, 5, 1)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    fn = 'uad_conv2d_wgrad' if kind == 'conv' else 'uad_convT2d_wgrad'
    res = {}
    for mode in (1, 0):
        call(fn, x.data_ptr(), dz.data_ptr(), dw.data_ptr(), B, H, H, Cin, Cout, 5, 0, mode, ws.data_ptr(), wsb, st())
        torch.cuda.synchronize()
        res[mode] = dw.double().cpu()
    # float64 reference on the GPU in chunks via the SIMT result is not independent: use torch conv in float64
    xd, dzd = x.double().permute(0, 3, 1, 2), dz.double().permute(0, 3, 1, 2)
    if kind == 'conv':
        xp = torch.nn.functional.pad(xd, (1, 2, 1, 2))
        ref = torch.nn.grad.conv2d_weight(xp, (Cout, Cin, 5, 5), dzd, stride=2).permute(2, 3, 1, 0)           # [kh, kw, Cin, Cout]
    else:
        # convT fwd: out = conv_transpose(x, w); its filter gradient = conv weight gradient with roles swapped: dW[kh,kw,co,ci]
        dzp = torch.nn.functional.pad(dzd, (1, 2, 1, 2))
        ref = torch.nn.grad.conv2d_weight(dzp, (Cin, Cout, 5, 5), xd, stride=2).permute(2, 3, 1, 0)            # [kh, kw, Cout, Cin]
    ref = ref.cpu()
    m = ref.abs().max()
    print(f"{name:12s} depth {os.environ.get('UAD_WS_DEPTH', 'default'):>7s} ss {os.environ.get('UAD_WGRAD_SS', '1')}: "
          f"tc rel-err {(res[1] - ref).abs().max() / m:.3e}   simt fp32 rel-err {(res[0] - ref).abs().max() / m:.3e}", flush=True)
