"""PeerDAS block of bench.py alone, for a few batch sizes: python tools/cells_bench.py [n ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import lambdaworks_kzg_b200 as lw  # noqa: E402

if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [256]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    for n in sizes:
        out = bench.cells_block(lw, torch, dev, n)
        print(json.dumps(out))
