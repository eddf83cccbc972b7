"""CPU, world_size 2 over gloo: the batch-sharded host logic (shard -> per-rank partial stats -> one
all-reduce -> global mean / dT / confusion matrix) reproduces the single-process answer.  The per-rank
partials come from the oracle here (no GPU); on the GPU box the fused kernel writes the same buffer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import class_dist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_stats(logits, T, labels, size):
    """What the kernel writes for one shard: {sum -log q, n_valid, raw dT} in float64."""
    from oracle import simt_oracle as O
    n_valid = int((labels != 255).sum())
    CK, C = T.shape
    if n_valid == 0:
        return torch.zeros(2 + CK * C, dtype=torch.float64)
    loss, _, dT = O.simt_head_fwd_bwd(logits, T, labels, size, torch.float64)
    return torch.cat([torch.stack([loss * n_valid, torch.tensor(float(n_valid), dtype=torch.float64)]),
                      (dT * n_valid).reshape(-1)])


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import simt_oracle as O
        from simt_b200 import dist as sd
        torch.set_num_threads(2)
        B, CK, C = 3, 23, 19            # 3 images over 2 ranks: ragged shards (2 + 1)
        logits, labels = O.synth_head_inputs(B, CK, 5, 9, 32, 64, seed=11, coherent=True, block=8,
                                             class_dist=class_dist())
        labels[2] = 255                 # the second rank's only image is fully ignored
        labels[1, :4] = 255
        torch.manual_seed(5)
        T = torch.softmax(torch.randn(CK, C, dtype=torch.float64), 1)
        lg = sd.shard_batch(logits, rank, world)
        lb = sd.shard_batch(labels, rank, world)
        stats = _oracle_stats(lg, T, lb, (32, 64))
        sd.reduce_head_stats(stats)
        loss, dT = sd.finish_head(stats, CK, C)
        # eval histogram, sharded over images
        gts, prs = zip(*[O.synth_eval_pair(48, 64, seed=20 + i, block=16) for i in range(5)])
        lo, hi = sd.shard_range(5, rank, world)
        hist = torch.zeros(19, 19, dtype=torch.int64)
        for i in range(lo, hi):
            lab = O.label_mapping(gts[i], np.array(O.CITYSCAPES_LABEL2TRAIN))
            hist += torch.from_numpy(O.fast_hist(lab.flatten(), prs[i].flatten(), 19))
        sd.reduce_hist(hist)
        torch.save({"loss": loss, "dT": dT, "hist": hist, "n": stats[1]}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_head_and_hist_equal_single_process(tmp_path):
    from oracle import simt_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    # every rank holds the same global answer
    assert torch.equal(res[0]["dT"], res[1]["dT"]) and torch.equal(res[0]["hist"], res[1]["hist"])
    B, CK, C = 3, 23, 19
    logits, labels = O.synth_head_inputs(B, CK, 5, 9, 32, 64, seed=11, coherent=True, block=8, class_dist=class_dist())
    labels[2] = 255
    labels[1, :4] = 255
    torch.manual_seed(5)
    T = torch.softmax(torch.randn(CK, C, dtype=torch.float64), 1)
    loss, _, dT = O.simt_head_fwd_bwd(logits, T, labels, (32, 64), torch.float64)
    assert abs(float(res[0]["loss"]) - float(loss)) <= 1e-12 * abs(float(loss))
    assert float((res[0]["dT"] - dT).norm() / dT.norm()) <= 1e-12
    assert int(res[0]["n"]) == int((labels != 255).sum())
    hist = np.zeros((19, 19), dtype=np.int64)
    for i in range(5):
        gt, pr = O.synth_eval_pair(48, 64, seed=20 + i, block=16)
        hist += O.fast_hist(O.label_mapping(gt, np.array(O.CITYSCAPES_LABEL2TRAIN)).flatten(), pr.flatten(), 19)
    assert np.array_equal(res[0]["hist"].numpy(), hist)            # bit-exact


def test_shard_range_partitions():
    from simt_b200.dist import shard_range
    for n in (0, 1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
