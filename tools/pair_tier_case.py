# development aid: ONE chain-bound batch whose plan is "pair tier for everything" (4 096 x 1 MiB SHA3-512: the warp-tier
# blocks would not all be resident), for an ncu capture of the pair tier inside the sponge
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import numpy as np, torch
from capycrypt_b200 import Engine
eng = Engine()
n = int(_sys.argv[1]) if len(_sys.argv) > 1 else 4096
lens = np.full(n, 1 << 20, dtype=np.int64)
off = np.zeros(n + 1, np.int64); off[1:] = np.cumsum(lens)
data = torch.empty(int(off[-1]) + 16, dtype=torch.uint8, device="cuda"); data.random_(0, 256)
t_off = torch.from_numpy(off).cuda()
out = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.sha3_dev(data, t_off, 512, out)
torch.cuda.synchronize()
print("done")
