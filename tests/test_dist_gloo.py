"""CPU, world_size 2 (and 3), gloo: the host logic of the bank-row-sharded mode —
partition, global index offsets, all-gather of candidate keys, merge — with the compute
steps served by the ORACLE (tests may use it; the product's default ops are CUDA-only).
Invariant: the G-shard result is bitwise the 1-shard result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import datagen
from oracle import knn_oracle as O


class OracleOps:
    """numpy/C oracle standing in for the CUDA library on CPU (same ops contract as _CudaOps)."""

    sample_stride = 0      # >0: emulate the sampling pre-pass with this row stride
    sample_r = 16
    poison_threshold = False  # force the (1e-7-probability) failure of the sampled bound

    @classmethod
    def topk_keys(cls, feature, bank_shard, k, mode, idx_offset, tau0=None):
        sims = O.sims_seqfma(feature.numpy(), bank_shard.numpy())
        if tau0 is not None:
            sims = np.where(sims > tau0.numpy()[:, None], sims, -np.inf)
        s, i = O.canonical_topk_c(sims, k, idx_offset)
        keys = O.make_keys(s, i)
        keys[np.isneginf(s)] = 0  # below the threshold: empty slot
        return torch.from_numpy(keys.view(np.int64))

    @classmethod
    def sample_keys(cls, feature, bank_shard, k, mode, n_rows_global):
        if cls.sample_stride == 0:
            return None
        sub = bank_shard.numpy()[:, :: cls.sample_stride]
        r = min(cls.sample_r, sub.shape[1])
        s, i = O.topk_seqfma(feature.numpy(), np.ascontiguousarray(sub), r)
        return torch.from_numpy(O.make_keys(s, i).view(np.int64))

    @classmethod
    def kth_sim(cls, keys):
        s, _ = O.decode_keys(keys.numpy().view(np.uint64))
        t = torch.from_numpy(np.ascontiguousarray(s[:, -1]))
        if cls.poison_threshold:
            t[0] = 10.0  # nothing passes for row 0 -> must be repaired
        return t

    @staticmethod
    def merge_keys(keys_in, k_out):
        return torch.from_numpy(O.merge_keys_np(keys_in.numpy().view(np.uint64), k_out).view(np.int64))

    @staticmethod
    def vote(keys, labels, num_classes, knn_t):
        s, i = O.decode_keys(keys.numpy().view(np.uint64))
        return torch.from_numpy(O.vote_o64(s, i, labels.numpy(), num_classes, knn_t)[0])

    @classmethod
    def vote_packed(cls, keys, labels, num_classes, knn_t, n_rows_out):
        out = torch.zeros((n_rows_out, num_classes + 1), dtype=torch.int64)
        n = keys.shape[0]
        if n:
            out[:n, :num_classes] = cls.vote(keys, labels, num_classes, knn_t)
            out[:n, num_classes] = (keys[:, -1] == 0).to(torch.int64)
        return out

    @staticmethod
    def decode_keys(keys):
        s, i = O.decode_keys(keys.numpy().view(np.uint64))
        return torch.from_numpy(s), torch.from_numpy(i)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir, stride=0, poison=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from b200knn import ShardedBank

        OracleOps.sample_stride, OracleOps.poison_threshold = stride, poison
        c = datagen.make_case(case)
        bank = torch.from_numpy(c["bank"])
        sb = ShardedBank.from_full(bank, torch.from_numpy(c["labels"]), ops=OracleOps, mode="bf16")
        q = torch.from_numpy(c["feature"])
        keys = sb.topk_keys(q, c["k"])
        pred = sb.knn_predict(q, c["C"], c["k"], c["t"])  # all-to-all by query slice (default)
        pred_ag = sb.knn_predict(q, c["C"], c["k"], c["t"], exchange="allgather")
        assert torch.equal(pred, pred_ag)
        np.save(os.path.join(out_dir, f"keys_{rank}.npy"), keys.numpy())
        np.save(os.path.join(out_dir, f"pred_{rank}.npy"), pred.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case,stride,poison", [
    (2, "clustered_small", 0, False),   # no pre-pass
    (3, "ragged", 0, False),
    (2, "k5", 0, False),
    (2, "clustered_small", 4, False),   # global sampled threshold, all-reduced across shards
    (3, "gauss_small", 4, True),        # ... with one row's threshold forced too high -> repaired
])
def test_sharded_equals_single(world, case, stride, poison, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path), stride, poison), nprocs=world, join=True)
    c = datagen.make_case(case)
    s, i = O.topk_seqfma(c["feature"], c["bank"], c["k"])
    want_keys = O.make_keys(s, i).view(np.int64)
    want_pred = O.vote_o64(s, i, c["labels"], c["C"], c["t"])[0]
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"keys_{r}.npy"), want_keys)  # bitwise, every rank
        assert np.array_equal(np.load(tmp_path / f"pred_{r}.npy"), want_pred)


def test_shard_smaller_than_k(tmp_path):
    """ragged: N=1001 over 3 ranks with k=10 is fine; here k exceeds one shard's rows."""
    from b200knn import ShardedBank

    c = datagen.make_case("ragged")
    bank = torch.from_numpy(c["bank"][:, :6].copy())
    sb = ShardedBank(bank, torch.from_numpy(c["labels"][:6]), 6, ops=OracleOps)
    keys = sb.topk_keys(torch.from_numpy(c["feature"]), 6)
    assert keys.shape == (7, 6)
    with pytest.raises(RuntimeError, match="out of range"):
        sb.topk_keys(torch.from_numpy(c["feature"]), 7)
