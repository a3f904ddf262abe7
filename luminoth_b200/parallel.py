"""Multi-GPU plumbing for the image-sharded (data-parallel) predict path.

Images are independent units (the reference processes one image at a time,
``luminoth/tasks.py:146-154``), so the batch shards across ranks with no
data-path collective inside the forward.  The only exchanges are
(1) a one-off broadcast of the packed weight arena from rank 0 and
(2) a per-step all-gather of fixed-size padded detection records.
Backend-agnostic (``nccl`` on GPUs, ``gloo`` in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_weights(weights, specs):
    """dict -> one flat float32 array in `specs` order (name, shape)."""
    parts = []
    for name, shape in specs:
        a = np.asarray(weights[name], np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError("variable '%s' has shape %s, expected %s" % (name, a.shape, tuple(shape)))
        parts.append(a.reshape(-1))
    return np.concatenate(parts) if parts else np.zeros((0,), np.float32)


def unpack_weights(flat, specs):
    out, off = {}, 0
    for name, shape in specs:
        n = int(np.prod(shape))
        out[name] = np.asarray(flat[off:off + n]).reshape(shape)
        off += n
    if off != len(flat):
        raise ValueError('weight arena size mismatch: %d vs %d' % (off, len(flat)))
    return out


def broadcast_weights(weights, specs, device, src=0):
    """Rank `src` passes the weight dict; every rank returns the same dict."""
    total = int(sum(int(np.prod(s)) for _, s in specs))
    if dist.get_rank() == src:
        flat = torch.from_numpy(pack_weights(weights, specs)).to(device)
    else:
        flat = torch.empty(total, dtype=torch.float32, device=device)
    dist.broadcast(flat, src=src)
    return unpack_weights(flat.cpu().numpy(), specs)


def record_width(kmax):
    return 1 + 6 * kmax


def pack_detections(boxes, scores, labels, counts, out=None):
    """(B,K,4),(B,K),(B,K),(B,) tensors -> (B, 1+6K) float32 records {count, boxes, scores, labels}."""
    B, K = scores.shape
    if out is None:
        out = torch.empty((B, record_width(K)), dtype=torch.float32, device=scores.device)
    out[:, 0] = counts.to(torch.float32)
    out[:, 1:1 + 4 * K] = boxes.reshape(B, 4 * K)
    out[:, 1 + 4 * K:1 + 5 * K] = scores
    out[:, 1 + 5 * K:] = labels.to(torch.float32)
    return out


def unpack_detections(rec, kmax):
    B = rec.shape[0]
    counts = rec[:, 0].to(torch.int32)
    boxes = rec[:, 1:1 + 4 * kmax].reshape(B, kmax, 4)
    scores = rec[:, 1 + 4 * kmax:1 + 5 * kmax]
    labels = rec[:, 1 + 5 * kmax:].to(torch.int32)
    return boxes, scores, labels, counts


def all_gather_detections(rec, out=None):
    """Every rank contributes (B, W) records; returns (world*B, W) in rank order."""
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    if dist.get_backend() == 'gloo':
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec)
        out.copy_(torch.cat(parts, 0))
    else:
        dist.all_gather_into_tensor(out, rec)
    return out
