#!/usr/bin/env python
"""Homology-scores permutation test on B200s.

Drop-in for the reference script experiments/homology-scores-permutation-test/permtest.py:
the same flags (:12-35), the same two output files (:176-187: results JSON with pval,
effect_size, n_permutations, n_less_than, errs; joblib pickle of the real and random differences),
the same per-permutation unit (difference_in_means, :68-80) and the same error accounting
(a failing permutation is printed and counted, :97-99,162-164).

Permutations come from one of three sources:
  -sf / -tf FILES   the reference's own mode: pickles dict(C, nodes) / dict(D, nodes) written by
                    gen_random_network.py, one pair per permutation;
  --n_permutations N --mode graphs    the same experiment without the 2 x N pickles: every
                    permutation draws a degree-preserving random graph pair
                    (gen_random_network.py:50-84) and factors it on the GPU;
  --n_permutations N --mode indices   BASELINE.json config 4: graphs fixed, the homolog list is
                    shuffled with a seeded generator before the landmarks are taken.
Launched under torchrun the permutations are dealt round-robin to the ranks (no data-path
collective; one all-gather of three doubles per permutation at the end) and rank 0 writes the files.
-j/--n_jobs is accepted for compatibility (the GPU ranks replace the joblib workers).
"""
import argparse
import json
import os
import sys

import joblib
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from munk_b200 import permtest, util  # noqa: E402
from munk_b200.io import get_logger  # noqa: E402


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    parser.add_argument('-s', '--source_edgelist', type=str, required=True)
    parser.add_argument('-t', '--target_edgelist', type=str, required=True)
    parser.add_argument('-hom', '--homolog_list', type=str, required=True)
    parser.add_argument('-nl', '--n_landmarks', type=int, required=False)
    parser.add_argument('-lam', '--lam', type=float, required=False)
    parser.add_argument('-sf', '--source_data_files', nargs='*', type=str, default=None)
    parser.add_argument('-tf', '--target_data_files', nargs='*', type=str, default=None)
    parser.add_argument('-o', '--output_file', type=str, required=True)
    parser.add_argument('-dof', '--diffs_output_file', type=str, required=True)
    parser.add_argument('-j', '--n_jobs', type=int, required=False, default=2)
    parser.add_argument('--n_permutations', type=int, default=0,
                        help='generate this many permutations on the GPU instead of reading -sf/-tf pickles')
    parser.add_argument('--mode', choices=['graphs', 'indices'], default='graphs')
    parser.add_argument('--seed', type=int, default=0)
    return parser.parse_args(argv)


def main(args):
    import networkx as nx
    import torch
    log = get_logger()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank, local = int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        import logging
        import torch.distributed as tdist
        tdist.init_process_group('nccl', device_id=torch.device('cuda', local))
        if rank != 0:
            log.setLevel(logging.WARNING)
    from_files = bool(args.source_data_files)
    if from_files:
        assert len(args.source_data_files) == len(args.target_data_files or [])     # permtest.py:115
    elif args.n_permutations <= 0:
        raise SystemExit('give either -sf/-tf pickles or --n_permutations N')

    log.info('Loading homologs list from %s', args.homolog_list)
    raw_homologs = util.read_homolog_list(args.homolog_list)
    log.info('Loading source edgelist from %s', args.source_edgelist)
    source_G = nx.read_edgelist(args.source_edgelist, encoding='ascii')
    log.info('Loading target edgelist from %s', args.target_edgelist)
    target_G = nx.read_edgelist(args.target_edgelist, encoding='ascii')
    log.info('Computing MUNK embeddings real PPI networks with %d landmarks', args.n_landmarks)
    source_G = util.simple_two_core(source_G, verbose=rank == 0)
    target_G = util.simple_two_core(target_G, verbose=rank == 0)
    homologs = util.homologs_in_graphs(source_G, target_G, raw_homologs)
    lam = 0.05 if args.lam is None else args.lam

    if from_files:
        res = permtest.run_permutation_test_from_files(
            source_G, target_G, homologs, args.n_landmarks, args.source_data_files, args.target_data_files,
            lam=lam, rank=rank, world=world)
    else:
        res = permtest.run_permutation_test(source_G, target_G, homologs, args.n_landmarks, args.n_permutations,
                                            lam=lam, mode=args.mode, rank=rank, world=world, seed=args.seed)
    if rank == 0:
        real, rand = res['real'], res['random']
        log.info('Difference between homolog and non-homolog similarity scores %f', real['mean_diff'])
        print('ERRORS:', res['errs'])
        log.info('Mean difference in means between homologs and non-homologs scores for random graphs %f',
                 np.mean(rand['mean_diffs']))
        log.info('N permutations less than real: %d, N permutations %d', res['n_less_than'], res['n_permutations'])
        log.info('P-value: %f', res['pval'])
        log.info('Effect-size: %f', res['effect_size'])
        log.info('# permutations: %d', len(rand['mean_diffs']) + res['errs'])
        results = dict(pval=res['pval'], effect_size=res['effect_size'], n_permutations=res['n_permutations'],
                       n_less_than=res['n_less_than'], errs=res['errs'])
        log.info('Writing results to %s', args.output_file)
        with open(args.output_file, 'w') as OUT:
            json.dump(results, OUT, indent=2)
        diffs = dict(real=dict(mean_diff=real['mean_diff'], non_hom_mean=real['non_hom_mean'],
                               hom_mean=real['hom_mean']),
                     random=dict(mean_diffs=tuple(float(x) for x in rand['mean_diffs']),
                                 non_hom_means=tuple(float(x) for x in rand['non_hom_means']),
                                 hom_means=tuple(float(x) for x in rand['hom_means'])))
        log.info('Writing values of differences to %s', args.diffs_output_file)
        joblib.dump(diffs, args.diffs_output_file)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == '__main__':
    main(parse_args())
