#!/usr/bin/env python
"""MUNK embeddings + similarity scores from two edge lists and a homolog list, on a B200.

Drop-in for the reference CLI (scripts/compute_embeddings.py:17-31 flag surface, :72-98 output
files and schemas).  Graph ingestion (parse, largest component, self loops, 2-core, sorted index
map) runs on arrays (munk_b200.ingest, bit-exact with the reference's networkx path; pass
--networkx_ingest to use networkx itself); the numeric work (both regularized-Laplacian kernels,
the source RKHS factor, the landmark projection and the n1 x n2 score product) runs in one
device-resident pass (munk_b200.pipeline.embed_and_score); the embeddings come down and their
pickles are written by worker threads while the score product is still running.  Launched under torchrun
(`python -m torch.distributed.run --nproc-per-node N scripts/compute_embeddings.py ...`) the same
pass is sharded over N GPUs and rank 0 writes the files.
"""
import argparse
import json
import os
import random
import sys
import time

import joblib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

import munk_b200 as munk  # noqa: E402
from munk_b200 import ingest, pipeline, util  # noqa: E402

REQUIRED = [
    ('-se', '--source_edgelist'), ('-te', '--target_edgelist'), ('-hf', '--homolog_list'),
    ('-so', '--source_output_file'), ('-to', '--target_output_file'),
    ('-sim', '--sim_scores_output_file'), ('-lo', '--landmarks_output_file'),
    ('-r', '--runtimes_file'),
]


def build_parser():
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    for short, long_name in REQUIRED:
        ap.add_argument(short, long_name, type=str, required=True)
    ap.add_argument('-n', '--n_landmarks', type=int, default=400)
    ap.add_argument('-rs', '--random_seed', type=int, default=28791)
    ap.add_argument('--src_lam', type=float, default=0.05)
    ap.add_argument('--tgt_lam', type=float, default=0.05)
    ap.add_argument('--networkx_ingest', action='store_true',
                    help='load and prune the graphs with networkx (the reference path) instead of arrays')
    return ap


def load_with_networkx(path, log, what):
    import networkx as nx
    log.info('Loading %s edgelist from %s', what, path)
    G = util.simple_two_core(nx.read_edgelist(path, encoding='ascii'))
    nodes = util.sorted_nodes(G)
    return nodes, util.edge_index_arrays(G, nodes)


def load_with_arrays(path, log, what):
    log.info('Loading %s edgelist from %s', what, path)
    try:
        nodes, u, v = ingest.load_two_core(path)
    except ingest.WeightedEdgeList:
        # nx.read_edgelist keeps edge weights and nx.laplacian_matrix honours them (munk.py:82):
        # weighted lists go through networkx, which returns the weighted Laplacian's entries
        log.info('%s edgelist is weighted: loading it with networkx', what)
        return load_with_networkx(path, log, what)
    log.info('2 core info - # Nodes: %d, # Edges: %d', len(nodes), len(u))
    return nodes, (u, v)


def run(args):
    log = munk.io.get_logger()
    random.seed(args.random_seed)   # accepted for compatibility; landmarks are the first k pairs

    # Multi-GPU is opt-in: launched under torchrun (WORLD_SIZE > 1) the pipeline is sharded over
    # the ranks (munk_b200.dist) and rank 0 writes the same files; a plain `python` launch is the
    # single-GPU drop-in.
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank, local = int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import logging
        import torch
        import torch.distributed as tdist
        from munk_b200 import dist as mdist
        from munk_b200.device import Device
        torch.cuda.set_device(local)
        tdist.init_process_group('nccl', device_id=torch.device('cuda', local))
        if rank != 0:
            log.setLevel(logging.WARNING)

    log.info('Loading homologs list from %s', args.homolog_list)
    raw_homologs = util.read_homolog_list(args.homolog_list)
    load = load_with_networkx if args.networkx_ingest else load_with_arrays
    source_nodes, source_edges = load(args.source_edgelist, log, 'source')
    target_nodes, target_edges = load(args.target_edgelist, log, 'target')
    homologs = ingest.homologs_in_node_lists(source_nodes, target_nodes, raw_homologs)

    log.info('Computing MUNK embeddings with %d landmarks', args.n_landmarks)
    inputs = pipeline.index_inputs(source_nodes, source_edges, target_nodes, target_edges, homologs,
                                   args.n_landmarks)
    # The three pickles are ~8 GB at 20k/16k nodes; written one after the other they take as long as
    # the whole compute stage.  They are written by worker threads (file writes release the GIL),
    # the two embeddings while the score product is still running on the GPU.
    from concurrent.futures import ThreadPoolExecutor
    writers = ThreadPoolExecutor(max_workers=3)
    pending = []

    def save_both(source_X, target_X):
        landmarks = inputs['landmark_homs']
        log.info('Saving source species embeddings to %s', args.source_output_file)
        pending.append(writers.submit(munk.save_embeddings, source_X, source_nodes, [a for a, _ in landmarks],
                                      args.source_output_file))
        log.info('Saving target species embeddings to %s', args.target_output_file)
        pending.append(writers.submit(munk.save_embeddings, target_X, target_nodes, [b for _, b in landmarks],
                                      args.target_output_file))

    started = time.time()
    if world > 1:
        res = mdist.embed_and_score_distributed(Device(local), inputs, rank, world,
                                                src_lam=args.src_lam, tgt_lam=args.tgt_lam)
        total_time = time.time() - started
        if res is not None:
            save_both(res['source_X'], res['target_X'])
    else:
        res = pipeline.embed_and_score(None, None, homologs, args.n_landmarks, src_lam=args.src_lam,
                                       tgt_lam=args.tgt_lam, inputs=inputs, on_embeddings=save_both)
        total_time = res['embed_wall']      # embed_networks up to host-resident embeddings (:60-65)
    if res is None:                 # ranks > 0 of a torchrun launch: rank 0 writes the files
        tdist.barrier()
        tdist.destroy_process_group()
        return
    landmarks = res['landmarks']
    log.info('Source data shape {}'.format(res['source_X'].shape))
    log.info('Target data shape {}'.format(res['target_X'].shape))

    log.info('Saving MUNK similarity scores to %s', args.sim_scores_output_file)
    pending.append(writers.submit(joblib.dump, dict(X=res['scores'], A_nodes=res['source_nodes'],
                                                    B_nodes=res['target_nodes'], landmarks=landmarks,
                                                    homologs=homologs), args.sim_scores_output_file))

    log.info('Saving landmark list to %s', args.landmarks_output_file)
    with open(args.landmarks_output_file, 'w') as out:
        out.writelines('%s\t%s\n' % pair for pair in landmarks)

    log.info('Saving runtimes to %s', args.runtimes_file)
    with open(args.runtimes_file, 'w') as out:
        record = dict(n_source_nodes=len(source_nodes), n_target_nodes=len(target_nodes),
                      runtimes=res['runtimes'], total_time=total_time)
        if res.get('timing_notes'):
            # beyond the reference's schema: how to read the four timers (phases overlap on the GPU)
            record['timing_notes'] = res['timing_notes']
        json.dump(record, out, indent=2)
    for job in pending:
        job.result()
    writers.shutdown()
    log.info('MUNK embedding complete!')
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == '__main__':
    run(build_parser().parse_args())
