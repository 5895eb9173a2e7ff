import sys
sys.path.insert(0, "/root/repo")
import torch, json
from streamvln_b200 import microbench
import bench
pk = bench.peaks()
r = microbench.bench_kv_attn(torch.device("cuda:0"), pk)
for row in r["rows"][:2]: print(row["kernel"], row["achieved"], row["frac"])
