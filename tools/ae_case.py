# development aid: the sponge-AE seal (sponge_kernel2) and open over 2^16 x 4 KB, for ncu captures
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import torch
from capycrypt_b200 import Engine
eng = Engine()
n, mlen = 1 << 16, 4096
g = torch.Generator(device="cuda"); g.manual_seed(1)
rnd = lambda k: torch.randint(0, 256, (k,), dtype=torch.uint8, device="cuda", generator=g)
data, pw, nonces = rnd(n * mlen), rnd(n * 32), rnd(n * 512)
po = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 32
mo = torch.arange(n + 1, dtype=torch.int64, device="cuda") * mlen
ct, tag = torch.empty_like(data), torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.sponge_encrypt_dev(pw, po, n * 32, nonces, 512, data, mo, 512, ct, tag)
torch.cuda.synchronize()
print("done")
