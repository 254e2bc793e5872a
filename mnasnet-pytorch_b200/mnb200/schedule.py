"""Shape-change scheduling around the training step.

The hot path is planned per input shape (N, H, W): activation buffers, bound launches and (single GPU) a CUDA
graph.  The reference changes that shape in two ways, both restated here as host logic so the plans can be built
before the first step of a new shape instead of inside it:

* progressive resizing (src/train.py:305-317): every `epochs_grow_size` epochs, while `size_ratio < 1`, the image
  side doubles (`size_ratio *= 2`) and the loader's batch size is divided by 4 (`batch_size // 4`);
* cluster-batched rectangular crops (src/utils/cluster_random_sampler.py:18-52, src/utils/datasets.py:331-335,440):
  images are grouped into aspect clusters {0: 384x512, 1: 512x512, 2: 512x384} (H x W), scaled by `size_ratio`;
  every batch is drawn from ONE cluster, incomplete batches are dropped, batches are shuffled.
"""
import random
from typing import Dict, Iterable, List, Sequence, Tuple

CLUSTER_DICT = {0: (384, 512), 1: (512, 512), 2: (512, 384)}      # src/utils/datasets.py:331-335, (H, W)


def cluster_shape(cluster: int, size_ratio: float, cluster_dict: Dict[int, Tuple[int, int]] = CLUSTER_DICT):
    """`[int(_ * self.size_ratio) for _ in target_size]` (src/utils/datasets.py:440)."""
    h, w = cluster_dict[cluster]
    return int(h * size_ratio), int(w * size_ratio)


def progressive_resize(epochs: int, epochs_grow_size: int, size_ratio: float, batch_size: int, start_epoch: int = 0):
    """[(epoch, size_ratio, batch_size)] as src/train.py:305-317 evolves them (the change applies from that epoch)."""
    out = []
    for epoch in range(start_epoch, epochs):
        if epochs_grow_size > 0 and (epoch + 1) % epochs_grow_size == 0 and size_ratio < 1.0:
            size_ratio = size_ratio * 2
            batch_size = int(batch_size // 4)
        out.append((epoch, size_ratio, batch_size))
    return out


def cluster_batches(cluster_indices: Sequence[Sequence[int]], batch_size: int, shuffle: bool = True,
                    oversampling: Sequence[Sequence[int]] = None, rng: random.Random = None):
    """Batches of sample indices, each from one cluster, as ClusterRandomSampler builds and iterates them
    (cluster_random_sampler.py:18-52): optional per-sample oversampling counts, chunks of `batch_size`, short chunks
    dropped, batch order shuffled (once at construction, once more per `__iter__`).  Returns [(cluster, [indices])]."""
    rng = rng or random
    lists = []
    for j, idx in enumerate(cluster_indices):
        idx = list(idx)
        if oversampling is not None:
            assert len(oversampling[j]) == len(idx)
            idx = [i for k, i in enumerate(idx) for _ in range(oversampling[j][k])]
            if shuffle:
                rng.shuffle(idx)
        batches = [idx[i:i + batch_size] for i in range(0, len(idx), batch_size)]
        batches = [(j, b) for b in batches if len(b) == batch_size]
        if shuffle:
            rng.shuffle(batches)
        lists.append(batches)
    flat = [b for lst in lists for b in lst]
    if shuffle:
        rng.shuffle(flat)           # construction
        rng.shuffle(flat)           # __iter__
    return flat


def step_shapes(batches: Iterable[Tuple[int, List[int]]], size_ratio: float, n_ranks: int = 1,
                cluster_dict: Dict[int, Tuple[int, int]] = CLUSTER_DICT):
    """Per-rank (N, H, W) of every step of an epoch: the global batch is split over `n_ranks` (batch sharding,
    DESIGN.md section 6), H x W comes from the batch's cluster."""
    out = []
    for cluster, idx in batches:
        h, w = cluster_shape(cluster, size_ratio, cluster_dict)
        out.append((len(idx) // n_ranks, h, w))
    return out


def distinct_shapes(schedule, clusters_present: Iterable[int], n_ranks: int = 1,
                    cluster_dict: Dict[int, Tuple[int, int]] = CLUSTER_DICT):
    """All (N, H, W) a run will need, in first-use order, from progressive_resize()'s output."""
    seen, out = set(), []
    for _, ratio, batch in schedule:
        for c in clusters_present:
            h, w = cluster_shape(c, ratio, cluster_dict)
            key = (batch // n_ranks, h, w)
            if key[0] > 0 and key not in seen:
                seen.add(key)
                out.append(key)
    return out


def warm_plans(engine, shapes: Iterable[Tuple[int, int, int]], graphs: bool = False, lr: float = 0.0):
    """Builds the plan (buffers + bound launches) of every shape ahead of time so the first step after a resize /
    cluster switch costs no allocation; with `graphs` the whole step is also captured (single GPU).  State is left
    untouched: graph capture restores parameters, BN buffers and optimizer state after its warm-up step."""
    import torch
    built = []
    for (n, h, w) in shapes:
        engine.plan(n, h, w)
        if graphs:
            x = torch.zeros(n, engine.in_channels, h, w, device=engine.device)
            t = torch.zeros(n, dtype=torch.long, device=engine.device)
            if (n, h, w) not in engine.graphs:
                # capture only: train_step_graph would also replay once, i.e. take a real optimizer step
                engine.capture_step_graph(x, t, lr)
        built.append((n, h, w))
    return built
