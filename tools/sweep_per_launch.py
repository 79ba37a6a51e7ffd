"""Config E (bench.sweep4096) on one GPU, alone: RXB_SWEEP_PER_LAUNCH=<frames per launch> and RXC_RASTER_GROUPS=<n> are the knobs.
usage: sweep_per_launch.py"""
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
print('per_launch', os.environ.get('RXB_SWEEP_PER_LAUNCH', 'default'), 'groups', os.environ.get('RXC_RASTER_GROUPS', 'default'), 'render-only s / delivered s / host prepare s:', r['render_only']['total_job_s'], r['delivered_to_rank0']['total_job_s'], r['host_prepare_s'], flush=True)
