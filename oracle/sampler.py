"""ORACLE (test infrastructure only -- never imported by the product path).

Bit-exact restatement of the integer sampling on the triplet path:

* torch's CPU generator (mt19937; `torch.randint` = next_u32 % range + low,
  `Tensor.random_()` on int64 = next_u64 % 2**63, `torch.randperm` Fisher-Yates)
  as consumed by dataset.py:76 (`torch.randint(0, taskNum)`) and
  Envs/audioLoader.py:174-176 (`rand_fn(0, len(...), size=())`);
* the negative-class rule dataset.py:72-78 (collision -> "empty" class taskNum);
* the per-item draw order of dataset.py:34-62 (`getImgSoundPair`);
* DataLoader(shuffle=True, num_workers=0) batch order (RandomSampler);
* the iTHOR variant: task list of dataset.py:17-29, `getAudioFromTask`
  (Envs/audioLoader.py:223-237: location-synonym draw, object-synonym draw) and
  `genSoundFeatFromTask` (Envs/audioLoader.py:203-209: clip draw) -- three draws per sound;
* `from_torch_state`: continue torch's global CPU generator (torch.get_rng_state()).

PINNED: oracle/make_golden.py drives the imported reference `VARDataset` /
`DataLoader` with a recording audio stub (Kuka) and with the reference's own `audioLoader`
methods under the real `AI2ThorConfig` + `EnvConfig` (iTHOR) under the same seeds and stores
the index streams in tests/golden/sampler_*.npz.
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


def torch_state_words(rng_state):
    """torch.get_rng_state() bytes (the legacy THGeneratorState layout at::CPUGeneratorImpl serialises:
    u64 the_initial_seed, i32 left, i32 seeded, u64 next, u64 state[624], normal-sample cache)
    -> (words[624] uint32, pos) with pos = 624 when the next draw twists first."""
    b = np.asarray(rng_state, dtype=np.uint8).tobytes()
    left = int(np.frombuffer(b, dtype=np.int32, count=1, offset=8)[0])
    nxt = int(np.frombuffer(b, dtype=np.uint64, count=1, offset=16)[0])
    words = np.frombuffer(b, dtype=np.uint64, count=_N, offset=24).astype(np.uint32)
    return words, (_N if left <= 1 else nxt)


class TorchCPUGenerator:
    """at::CPUGeneratorImpl draws as used by the reference's sampling calls."""

    def __init__(self, seed):
        self.engine = MT19937(seed)

    @classmethod
    def from_torch_state(cls, rng_state):
        g = cls(0)
        g.engine.state, g.engine.pos = torch_state_words(rng_state)
        return g

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


def ithor_task_tables(all_tasks, synonym, obj_act, words):
    """Flatten config.allTasks / config.synonym / soundSource['FSC_obj_act'] and the loaded
    words[loc][obj][act] clip lists into the tables the device sampler indexes.
    -> (n_loc[t], n_obj[t], lists[t][li][oi] = (fsc_loc, fsc_obj, fsc_act)) with the task order of
    dataset.py:22-29 and the action resolved as Envs/audioLoader.py:230-232."""
    n_loc, n_obj, lists = [], [], []
    for loc in all_tasks:
        for obj in all_tasks[loc]:
            for act in all_tasks[loc][obj]:
                ls, os_ = synonym[loc], synonym[obj]
                n_loc.append(len(ls)); n_obj.append(len(os_))
                row = []
                for fl in ls:
                    col = []
                    for fo in os_:
                        fa = sorted(set(obj_act[fo]).intersection(synonym[act]))
                        assert len(fa) == 1, "the reference takes element [0] of a set: only one match is deterministic"
                        assert fa[0] in words[fl][fo]
                        col.append((fl, fo, fa[0]))
                    row.append(col)
                lists.append(row)
    return n_loc, n_obj, lists


def draw_clip_ithor(gen, task, n_loc, n_obj, sizes):
    """getAudioFromTask + genSoundFeatFromTask: sizes[task][li][oi] = clips in the resolved list."""
    li = gen.randint(0, n_loc[task])
    oi = gen.randint(0, n_obj[task])
    clip = gen.randint(0, sizes[task][li][oi])
    return task, li, oi, clip


def sample_triplet_ithor(gen, gt, task_num, n_loc, n_obj, sizes, stored_sn_id=None):
    """dataset.py:64-89 + :34-62 for config.name == 'AI2ThorConfig'."""
    sn_id = negative_class(gen, gt, task_num, stored_sn_id)
    if gt == task_num:
        pos = None
        neg = draw_clip_ithor(gen, sn_id, n_loc, n_obj, sizes)
    else:
        pos = draw_clip_ithor(gen, gt, n_loc, n_obj, sizes)
        neg = None if sn_id == task_num else draw_clip_ithor(gen, sn_id, n_loc, n_obj, sizes)
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
