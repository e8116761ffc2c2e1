import sys, torch, numpy as np
sys.path.insert(0, '.')
from optimistix_b200 import _lib
lib = _lib.load(); N = 256; DEV = 'cuda:0'
m = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
mode = sys.argv[2] if len(sys.argv) > 2 else 'rand'
if mode == 'rand':
    J = torch.randn(m, N, device=DEV)
elif mode == 'col':   # J[r][c] = c+1 on row 0 only
    J = torch.zeros(m, N, device=DEV); J[0] = torch.arange(1, N + 1, device=DEV).float()
elif mode == 'eye':
    J = torch.zeros(m, N, device=DEV); idx = torch.arange(min(m, N)); J[idx, idx] = 1.0
r = torch.randn(m, device=DEV)
ws = torch.empty(lib.b200opt_gn_gram_workspace_size(N), dtype=torch.uint8, device=DEV)
out = torch.full((N * N + N + 1,), float('nan'), dtype=torch.float32, device=DEV)
rc = lib.b200opt_gn_gram(J.data_ptr(), r.data_ptr(), m, N, out.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print('rc', rc, lib.b200opt_last_error())
G = out[:N * N].view(N, N).double().cpu(); g = out[N * N:N * N + N].double().cpu(); rr = out[-1].item()
Jd = J.double().cpu(); rd = r.double().cpu()
Gr = Jd.T @ Jd; gr = Jd.T @ rd
print('rr', rr, float(rd @ rd))
print('g err', float((g - gr).abs().max()), 'g[:4]', g[:4].tolist(), gr[:4].tolist())
print('G nan', int(torch.isnan(G).sum()), 'G zero frac', float((G == 0).double().mean()))
np.set_printoptions(precision=3, suppress=True, linewidth=200)
print('G[:6,:6]\n', G[:6, :6].numpy()); print('ref\n', Gr[:6, :6].numpy())
for (a0, b0) in [(0, 0), (0, 128), (128, 128), (128, 0), (0, 32), (32, 0), (64, 200)]:
    blk = (G[a0:a0 + 32, b0:b0 + 32] - Gr[a0:a0 + 32, b0:b0 + 32]).abs().max() / Gr.abs().max()
    print('block', a0, b0, 'rel err', float(blk))
