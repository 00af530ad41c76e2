"""torchrun check of the N>1 scan path: every rank runs perform_scan(), rows are all-gathered over NCCL; rank 0
compares with the reference's 5x5 scan fixture.   torchrun --nproc-per-node 2 tools/dist_scan_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import acropolis_b200.flags as flags
flags.verbose = False
from acropolis_b200.models import DecayModel
from acropolis_b200.scans import BufferedScanner, ScanParameter
rows = BufferedScanner(DecayModel, mphi=ScanParameter(0, 3, 5), tau=ScanParameter(3, 10, 5), temp0=10., n0a=1e-10,
                       bree=0., braa=1.).perform_scan()
ref = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scan_decay_5x5.npz"))["rows"]
big = np.abs(ref) > 1e-30
err = np.max(np.abs(rows[big] / ref[big] - 1))
print("rank %d/%d: rows %s, max rel err vs reference %.2e" % (dist.get_rank(), dist.get_world_size(), rows.shape, err), flush=True)
assert rows.shape == ref.shape and err < 1e-4
rows_f = BufferedScanner(DecayModel, mphi=ScanParameter(0.1, 1.6, 4), tau=1e7, temp0=10., n0a=ScanParameter(-14, -3, 8, fast=True),
                         bree=0., braa=1.).perform_scan()
reff = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scan_fast.npz"))["decay_rows"]
assert rows_f.shape == reff.shape and np.array_equal(rows_f[:, :2], reff[:, :2])
dist.barrier()
dist.destroy_process_group()
