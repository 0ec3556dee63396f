"""CPU check of the item order of the packet kernel's item mode (csrc/rt_kernels.cu k_packet, "Item order"): the map
item -> (block, sample) -- 8 different neighbouring blocks per chunk at one sample, the next chunk the same blocks at the
next sample -- restated in numpy must hit every (block, sample) pair exactly once for any block count (ragged last group)
and batch size, and every full chunk must hold 8 DIFFERENT blocks at ONE sample."""
import numpy as np
import pytest

K_CHUNK = 8          # rt_kernel_common.cuh kChunk


def item_to_block_sample(item, n_work, batch):
    per_group = K_CHUNK * batch
    g = item // per_group
    r = item - g * per_group
    gs = np.minimum(K_CHUNK, n_work - g * K_CHUNK)
    return g * K_CHUNK + r % gs, r // gs


@pytest.mark.parametrize("n_work", [1, 7, 8, 9, 15, 16, 17, 255, 1013, 8100, 64800])
@pytest.mark.parametrize("batch", [1, 2, 3, 8, 16])
def test_item_order_is_a_bijection(n_work, batch):
    items = np.arange(n_work * batch, dtype=np.int64)
    w, sb = item_to_block_sample(items, n_work, batch)
    assert w.min() >= 0 and w.max() == n_work - 1 and sb.min() >= 0 and sb.max() == batch - 1
    assert len(np.unique(w * batch + sb)) == n_work * batch
    full_groups = n_work // K_CHUNK
    if full_groups:
        n_full = full_groups * K_CHUNK * batch
        wc = w[:n_full].reshape(-1, K_CHUNK)
        sc = sb[:n_full].reshape(-1, K_CHUNK)
        assert (np.sort(wc, axis=1) == wc[:, :1] // K_CHUNK * K_CHUNK + np.arange(K_CHUNK)).all()      # 8 different neighbouring blocks
        assert (sc == sc[:, :1]).all()                                                                  # at one sample
