"""2-GPU check of the optional latency mode: Ref-sharded search == single-GPU search, bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'c2-matching_b200'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import seeding
from c2m_b200 import ops
from c2m_b200.dist import corr_argmax_ref_sharded
rank, lr = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
fin = torch.stack([seeding.unit_features(7, 256, 60, 64)]).cuda()
fref = torch.stack([seeding.unit_features(8, 256, 90, 70)]).cuda()
idx_s, val_s = corr_argmax_ref_sharded(fin, fref, norm_input=True)
idx_1, val_1 = ops.corr_argmax(fin, fref, norm_input=True)
same = torch.equal(idx_s, idx_1) and torch.equal(val_s, val_1)
t = torch.tensor([int(same)], device='cuda')
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print('ref-sharded == single GPU:', bool(t.item()), 'world', dist.get_world_size())
dist.destroy_process_group()
