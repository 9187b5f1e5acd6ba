"""Lane-level model of label_fix_kernel's index queue (recboard_b200/csrc/simt.cuh): the owner's warp scans the
label vector 32 rows at a time, queues the matching row indices and flushes 32 at a time, carrying the remainder
over.  The model must visit exactly the matching rows in ascending order (the summation order the deterministic
one-hot correction promises) for any placement of the matches; the kernel's arithmetic is checked on the GPU
(tests/test_gpu_parity.py::test_ce_dw_scattered_hot_label)."""
import random


def visit_order(labels, i):
    m, mine = len(labels), labels[i]
    q, pend, order = [None] * 64, 0, []
    for jb in range(i & ~31, m, 32):
        mask = [i <= jb + lane < m and labels[jb + lane] == mine for lane in range(32)]
        for lane in range(32):
            if mask[lane]:
                slot = pend + sum(mask[:lane])          # __popc(hit & lanemask_lt)
                assert slot < 64
                q[slot] = jb + lane
        pend += sum(mask)
        last = jb + 32 >= m
        while pend >= 32 or (last and pend > 0):
            n = min(pend, 32)
            order += q[:n]                              # flush: rows added in queue order
            rest = [q[lane + 32] if lane + 32 < pend else 0 for lane in range(32)]
            for lane in range(32):
                if lane + 32 < pend:
                    q[lane] = rest[lane]
            pend -= n
    assert pend == 0
    return order


def test_queue_visits_matches_in_index_order():
    rng = random.Random(1)
    for _ in range(2000):
        m = rng.choice([1, 5, 31, 32, 33, 64, 100, 257, 1000])
        n_labels = rng.choice([1, 2, 3, 10])
        labels = [rng.randrange(n_labels) for _ in range(m)]
        i = labels.index(labels[rng.randrange(m)])      # an owner: first row with its label
        assert visit_order(labels, i) == [j for j in range(m) if labels[j] == labels[i]]
