"""Dev probe: v2p_gzip_files on a synthetic protein FASTA image resident in HBM (run under ncu for per-kernel numbers)."""
import sys
import zlib

import numpy as np
import torch

sys.path.insert(0, ".")
from vcf2prot_b200.gzipdev import DeviceGzip

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(1)
aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", np.uint8)
img = aa[rng.integers(0, 20, mb << 20)]
pos = np.sort(rng.integers(0, len(img), (mb << 20) // 550))
img[pos] = 10
img[np.minimum(pos + 1, len(img) - 1)] = ord(">")
fb = np.linspace(0, len(img), n_files + 1).astype(np.uint64)
dev = torch.device("cuda", 0)
d_img = torch.from_numpy(img).to(dev)
gz = DeviceGzip(0)
cap = gz.bound(len(img), n_files)
d_gz = torch.empty(cap, dtype=torch.uint8, device=dev)
for i in range(3):
    ob, res = gz.compress_device(d_img.data_ptr(), fb, d_gz.data_ptr(), cap)
    print("run", i, "ms", round(res.ms, 3), "GB/s", round(res.in_bytes / res.ms / 1e6, 1), "ratio", round(res.in_bytes / res.out_bytes, 3))
h = d_gz[: int(ob[-1])].cpu().numpy()
d = zlib.decompressobj(wbits=31)
assert d.decompress(h[int(ob[0]):int(ob[1])].tobytes()) == img[int(fb[0]):int(fb[1])].tobytes() and d.eof
print("ok")
