import sys, json
sys.path.insert(0, '/root/repo')
import torch
from finitevolumesolver_b200 import capi
import bench_secondary as bs
torch.cuda.set_device(0)
print(json.dumps(bs.amr_record(capi, torch, 6536.7, n0=int(sys.argv[1]), steps=int(sys.argv[2]))))
print(json.dumps(bs.eb_record(capi, torch, 6536.7, n=(int(sys.argv[3]), int(sys.argv[4])), steps=int(sys.argv[2]))))
