"""Entry point with the reference's CLI (main.py:40-61: -m -b -e -mlr -w -lf -pm, plus its unused
-t -milr -wd -snt), additive flags --dtype / --synthetic / --cuda-graph / --max-iters.

    python -m jck_generation_b200.main -m DCGAN -b 512 -e 1 -mlr 2e-4 --synthetic 1
    torchrun --nproc-per-node 8 -m jck_generation_b200.main -m DCGAN -b 512 ...     (data parallel)
"""
import argparse
import os
import random
from datetime import datetime

import numpy as np
import torch

from .change_randomseed import RANDOMSEED
from .enums import ModelEnum
from .logger.main_logger import MainLogger


def seed_everything(seed=RANDOMSEED):
    """reference main.py:31-37"""
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def get_arg_parse(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument('-t', '--test', type=int, default=0)
    p.add_argument('-pm', '--model_path', type=str, default='')
    p.add_argument('-lf', '--log_file', type=int, default=1)
    p.add_argument('-m', '--model', type=ModelEnum, choices=list(ModelEnum), default=ModelEnum.DCGAN)
    p.add_argument('-w', '--num_worker', type=int, default=0)
    p.add_argument('-b', '--batch_size', type=int, default=128)
    p.add_argument('-e', '--epoch', type=int, default=100)
    p.add_argument('-mlr', '--max_learning_rate', type=float, default=0.1)
    p.add_argument('-milr', '--min_learning_rate', type=float, default=1e-4)
    p.add_argument('-wd', '--weight_decay', type=float, default=5e-4)
    p.add_argument('-snt', '--nesterov', type=int, default=1)
    # additive
    p.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    p.add_argument('--synthetic', type=int, default=0)
    p.add_argument('--synthetic_batches', type=int, default=391)
    p.add_argument('--cuda-graph', dest='cuda_graph', type=int, default=1)
    p.add_argument('--max-iters', dest='max_iters', type=int, default=0)
    p.add_argument('--metrics', type=int, default=1)
    return p.parse_args(argv)


def main(args):
    datetime_now = args.model_path if args.model_path != '' else datetime.now().strftime("%Y%m%d_%H%M%S")
    args.save_path = os.path.join('.', 'save', str(args.model).lower(), datetime_now)
    os.makedirs(args.save_path, exist_ok=True)
    logger = MainLogger(args)
    logger.debug(f'args: {vars(args)}')
    logger.debug('init data preprocessing')
    if args.model == ModelEnum.DCGAN:
        from .model import DCGAN
        from .preprocess.dcgan_data_preprocessor import DCGANDataPreprocessor
        from .train.dcgan_trainer import DCGANTrainer
        data_pre = DCGANDataPreprocessor(args)
        data_pre.transform_data()
        trainer = DCGANTrainer(args, DCGAN.Generator(), DCGAN.Discriminator(), data_pre)
    else:
        from .model import CGAN
        from .preprocess.cgan_data_preprocessor import CGANDataPreprocessor
        from .train.cgan_trainer import CGANTrainer
        data_pre = CGANDataPreprocessor(args)
        data_pre.transform_data()
        trainer = CGANTrainer(args, CGAN.Generator(), CGAN.Discriminator(), data_pre)
    trainer.train()


if __name__ == "__main__":
    seed_everything()
    main(get_arg_parse())
