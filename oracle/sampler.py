"""ORACLE (test infrastructure only -- never imported by the product path).

Bit-exact restatement of the integer sampling on the triplet path:

* torch's CPU generator (mt19937; `torch.randint` = next_u32 % range + low,
  `Tensor.random_()` on int64 = next_u64 % 2**63, `torch.randperm` Fisher-Yates)
  as consumed by dataset.py:76 (`torch.randint(0, taskNum)`) and
  Envs/audioLoader.py:174-176 (`rand_fn(0, len(...), size=())`);
* the negative-class rule dataset.py:72-78 (collision -> "empty" class taskNum);
* the per-item draw order of dataset.py:34-62 (`getImgSoundPair`);
* DataLoader(shuffle=True, num_workers=0) batch order (RandomSampler).

PINNED: oracle/make_golden.py drives the imported reference `VARDataset` /
`DataLoader` with a recording audio stub under the same seeds and stores the
index streams in tests/golden/sampler_*.npz.
"""
import numpy as np

_N, _M = 624, 397
_UPPER, _LOWER = 0x80000000, 0x7FFFFFFF
_MATRIX_A = 0x9908B0DF


class MT19937:
    """std::mt19937 / at::mt19937 (32-bit seeding via init_genrand)."""

    def __init__(self, seed):
        seed = int(seed) & 0xFFFFFFFF
        st = np.zeros(_N, dtype=np.uint64)
        st[0] = seed
        for i in range(1, _N):
            prev = int(st[i - 1])
            st[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & 0xFFFFFFFF
        self.state = st.astype(np.uint32)
        self.pos = _N

    def _twist(self):
        s = self.state.astype(np.uint64)
        for kk in range(_N):
            y = (int(s[kk]) & _UPPER) | (int(s[(kk + 1) % _N]) & _LOWER)
            v = int(s[(kk + _M) % _N]) ^ (y >> 1)
            if y & 1:
                v ^= _MATRIX_A
            s[kk] = v
        self.state = s.astype(np.uint32)
        self.pos = 0

    def next_u32(self):
        if self.pos >= _N:
            self._twist()
        y = int(self.state[self.pos])
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


class TorchCPUGenerator:
    """at::CPUGeneratorImpl draws as used by the reference's sampling calls."""

    def __init__(self, seed):
        self.engine = MT19937(seed)

    def random(self):
        return self.engine.next_u32()

    def random64(self):
        hi = self.engine.next_u32()
        lo = self.engine.next_u32()
        return (hi << 32) | lo

    def randint(self, low, high):
        """torch.randint(low, high, size=()) for ranges < 2**32."""
        return low + self.random() % (high - low)

    def random_int64(self):
        """torch.empty((), dtype=int64).random_()."""
        return self.random64() % (1 << 63)

    def randperm(self, n):
        """torch.randperm(n) on CPU (small-n path)."""
        r = list(range(n))
        for i in range(n - 1):
            z = self.random() % (n - i)
            r[i], r[z + i] = r[z + i], r[i]
        return r


def negative_class(gen, gt, task_num, stored_sn_id=None):
    """dataset.py:72-78."""
    if stored_sn_id is not None:
        return int(stored_sn_id)
    sn = gen.randint(0, task_num)
    return task_num if sn == gt else sn


def draw_clip_kuka(gen, intent, task_num, dataset_sizes):
    """Envs/audioLoader.py:166-177: clamp intent, draw dataset then clip.
    dataset_sizes[intent] = list of clip counts per loaded dataset (dict order)."""
    if intent > task_num - 1:
        intent = task_num - 1
    sizes = dataset_sizes[intent]
    ds = gen.randint(0, len(sizes))
    clip = gen.randint(0, sizes[ds])
    return intent, ds, clip


def sample_triplet_kuka(gen, gt, task_num, dataset_sizes, stored_sn_id=None):
    """dataset.py:64-89 + :34-62 for the pybullet config.  Returns
    (sn_id, pos, neg) with pos/neg = (intent, dataset, clip) or None for the all-zero feature."""
    sn_id = negative_class(gen, gt, task_num, stored_sn_id)
    if gt == task_num:
        pos = None
        neg = draw_clip_kuka(gen, sn_id, task_num, dataset_sizes)
    else:
        pos = draw_clip_kuka(gen, gt, task_num, dataset_sizes)
        neg = None if sn_id == task_num else draw_clip_kuka(gen, sn_id, task_num, dataset_sizes)
    return sn_id, pos, neg


def epoch_batches(gen, n_items, batch_size, drop_last=False):
    """Index batches of one `for ... in DataLoader(shuffle=True, num_workers=0)` pass,
    consuming the global generator exactly like torch: base_seed draw at iterator
    creation, sampler-seed draw at first next(), randperm from a fresh generator."""
    gen.random_int64()                 # _BaseDataLoaderIter._base_seed
    sampler_seed = gen.random_int64()  # RandomSampler: seed for its private generator
    perm = TorchCPUGenerator(sampler_seed).randperm(n_items)
    batches = [perm[i:i + batch_size] for i in range(0, n_items, batch_size)]
    if drop_last and batches and len(batches[-1]) < batch_size:
        batches.pop()
    return batches
