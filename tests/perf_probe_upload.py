import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from amplicon_sorter_b200 import synth
from amplicon_sorter_b200.engine import Engine
reads, _, _ = synth.make_config(5, scale=1.0)
buf, offs = synth.pack_reads(reads)
pb = torch.from_numpy(buf).pin_memory(); po = torch.from_numpy(offs.view(np.int64)).pin_memory()
hb, ho = pb.numpy(), po.numpy().view(np.uint64)
eng = Engine(0)
for rep in range(5):
    t0 = time.perf_counter(); eng.upload_reads(hb, ho); t1 = time.perf_counter()
    t2 = time.perf_counter(); eng.upload_reads(buf, offs); t3 = time.perf_counter()
    print(f"upload pinned {1e3*(t1-t0):.1f} ms   pageable {1e3*(t3-t2):.1f} ms", flush=True)
