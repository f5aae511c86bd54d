"""Inference CLI with the reference's interface (mmsr/test.py):

    python mmsr/test.py -opt options/test/test_C2_matching_mse.yml [--launcher none|pytorch]

`--launcher pytorch` (under torchrun) shards the pair list over one process per GPU with NCCL
and all-gathers the metrics; the reference's distributed validation does not work
(sr_model.py:160-162)."""
import argparse
import logging
import os
import os.path as osp
import sys
import time

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))

import torch  # noqa: E402

from mmsr.data import create_dataloader, create_dataset  # noqa: E402
from mmsr.models import create_model  # noqa: E402
from mmsr.utils.options import dict2str, dict_to_nonedict, parse  # noqa: E402


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('-opt', type=str, required=True, help='Path to option YAML file.')
    parser.add_argument('--launcher', choices=['none', 'pytorch', 'slurm'], default='none', help='job launcher')
    parser.add_argument('--local_rank', type=int, default=0)
    args = parser.parse_args(argv)
    opt = parse(args.opt, is_train=False)

    if args.launcher == 'none':
        opt['dist'] = False
        print('Disabled distributed testing.', flush=True)
        if torch.cuda.is_available():
            ids = opt.get('gpu_ids') or [0]
            torch.cuda.set_device(int(ids[0]) % torch.cuda.device_count())
    else:
        opt['dist'] = True
        local_rank = int(os.environ.get('LOCAL_RANK', args.local_rank))
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group(backend='nccl')
    opt = dict_to_nonedict(opt)

    os.makedirs(opt['path']['log'], exist_ok=True)
    logger = logging.getLogger('base')
    if not logger.handlers:
        logger.setLevel(logging.INFO)
        rank = torch.distributed.get_rank() if opt['dist'] else 0
        if rank == 0:
            logger.addHandler(logging.StreamHandler())
            logger.addHandler(logging.FileHandler(osp.join(
                opt['path']['log'], f"test_{opt['name']}_{time.strftime('%Y%m%d_%H%M%S')}.log")))
    logger.info(dict2str(opt))

    loaders = []
    for phase, dataset_opt in sorted(opt['datasets'].items()):
        test_set = create_dataset(dataset_opt)
        loaders.append(create_dataloader(test_set, dataset_opt, dist=opt['dist']))     # rank::world shard of the pair list
        logger.info(f"Number of test images in {dataset_opt['name']}: {len(test_set)}")
    model = create_model(opt)
    for loader in loaders:
        logger.info(f"Testing {loader.dataset.opt['name']}...")
        model.validation(loader, current_iter=opt['name'], tb_logger=None, save_img=bool(opt['save_img']))
    if opt['dist']:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
