import os, sys, json
sys.path.insert(0, '/root/repo')
import torch, bench
from rusterix_b200 import DeviceContext
dev = torch.device('cuda', 0)
ctx = DeviceContext.get(0)
ctx.set_vm_jit(2)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
def barrier(): torch.cuda.synchronize()
r = bench.sweep4096(0, 1, 0, dev, barrier, None)
print(os.environ.get('RXB_SWEEP_PER_LAUNCH'), r['render_only']['total_job_s'], r['delivered_to_rank0']['total_job_s'], r['host_prepare_s'], flush=True)
