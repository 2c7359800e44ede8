from typing import Dict, List, Optional

import numpy as np

from .experience import Experience


class ReplayBuffer:
    """FIFO transition store with the reference's interface (ref: replay_buffer.py:9-74): ``add_experience`` appends
    the flattened transitions and drops the oldest beyond ``buffer_size``; ``sample_minibatch`` draws indices with
    ``np.random.randint`` (global numpy RNG, with replacement) -- the same random stream and the same logical order
    (index 0 = oldest kept transition) as the reference.

    Storage is a structure-of-arrays ring: one contiguous float32 array per column, allocated
    on the first ``add_experience`` and grown geometrically up to ``buffer_size`` rows, so appending is a slice copy
    and a minibatch is ONE fancy-index gather per column instead of the reference's per-transition Python-list walk
    (1.2 ms per 256-sample minibatch there).  ``sample_indices`` + ``gather`` expose the two halves so the
    off-policy trainer can draw all S minibatches of a ``train`` call at once.  The reference's list attributes
    (``observations`` ...) remain available as read-only views in logical order."""

    COLUMNS = ("observations", "actions", "rewards", "next_observations", "dones")

    def __init__(self, buffer_size: int = int(1e6)) -> None:
        self.buffer_size = int(buffer_size)
        self.current_size: int = 0
        self._head: int = 0           # physical row of logical index 0
        self._capacity: int = 0       # allocated rows (<= buffer_size)
        self._cols: Dict[str, Optional[np.ndarray]] = {k: None for k in self.COLUMNS}
        self._dev = None              # device mirror (torch CUDA tensors, float32), built lazily by device_columns()
        self._dev_dirty: List = []    # physical row ranges written since the mirror was last refreshed

    # ---- storage ----
    def _allocate(self, rows: int, samples) -> None:
        rows = min(max(rows, 1024), self.buffer_size)
        for k, v in zip(self.COLUMNS, samples):
            first = np.asarray(v[0])
            # rewards stay float64 and dones bool, like the values the reference's lists hold
            dt = np.float64 if k == "rewards" else (np.bool_ if k == "dones" else np.float32)
            self._cols[k] = np.empty((rows,) + first.shape, dtype=dt)
        self._capacity = rows

    def _grow(self, need: int) -> None:
        new_cap = self._capacity
        while new_cap < need:
            new_cap *= 2
        new_cap = min(new_cap, self.buffer_size)
        order = self._physical(np.arange(self.current_size))
        for k in self.COLUMNS:
            old = self._cols[k]
            new = np.empty((new_cap,) + old.shape[1:], dtype=old.dtype)
            new[:self.current_size] = old[order]
            self._cols[k] = new
        self._head, self._capacity = 0, new_cap
        self._dev = None  # reallocated: the mirror is rebuilt on the next device_columns()

    def _physical(self, logical: np.ndarray) -> np.ndarray:
        return (self._head + logical) % self._capacity if self._capacity else logical

    def add_experience(self, experience: Experience) -> None:
        if hasattr(experience, "transition_columns"):  # PackedExperience: contiguous columns, no per-row Python objects
            new = experience.transition_columns()
        else:
            new = (experience.flattened_observations, experience.flattened_actions, experience.flattened_rewards,
                   experience.flattened_next_observations, experience.flattened_dones)
        n = len(new[0])
        if n == 0:
            return
        if self._capacity == 0:
            self._allocate(n, new)
        if n >= self.buffer_size:  # only the newest buffer_size transitions survive (front deletion in the reference)
            new = tuple(v[n - self.buffer_size:] for v in new)
            n = self.buffer_size
        total = self.current_size + n
        if min(total, self.buffer_size) > self._capacity:
            self._grow(min(total, self.buffer_size))
        overflow = max(0, total - self.buffer_size)  # oldest rows dropped
        self._head = (self._head + overflow) % self._capacity
        self.current_size -= overflow
        start = (self._head + self.current_size) % self._capacity
        first = min(n, self._capacity - start)
        for k, v in zip(self.COLUMNS, new):
            col = self._cols[k]
            arr = np.asarray(v, dtype=col.dtype).reshape((n,) + col.shape[1:])
            col[start:start + first] = arr[:first]
            if first < n:
                col[:n - first] = arr[first:]
        self.current_size += n
        if self._dev is not None:  # no mirror yet: its first build uploads everything anyway
            self._dev_dirty.append((start, first))
            if first < n:
                self._dev_dirty.append((0, n - first))
            if len(self._dev_dirty) > 256:  # many small appends between two train() calls: one full refresh is cheaper
                self._dev_dirty = self._live_ranges()

    def _live_ranges(self) -> List:
        first = min(self.current_size, self._capacity - self._head)  # the rows that hold data, wrap-around aware
        return [(self._head, first)] + ([(0, self.current_size - first)] if first < self.current_size else [])

    # ---- device mirror: the columns as float32 CUDA tensors, refreshed incrementally ----
    def physical_rows(self, logical: np.ndarray) -> np.ndarray:
        return self._physical(np.asarray(logical)).astype(np.int64)

    def ring(self):
        """(physical row of logical index 0, live rows, allocated rows): what a device-side index draw needs."""
        return self._head, self.current_size, self._capacity

    def device_columns(self):
        """(obs, act, rew, next_obs, done) float32 CUDA tensors with ``capacity`` rows each, in PHYSICAL row order
        (use ``physical_rows`` on the logical indices).  rewards float64 -> float32 and dones bool -> 0/1 are the casts
        the reference applies to every minibatch (td3.py:226-228).  Only rows written since the last call are uploaded."""
        import torch
        if self._capacity == 0:
            raise ValueError("device_columns: the buffer is empty")
        if self._dev is None:
            self._dev = tuple(torch.zeros(self._cols[k].shape, dtype=torch.float32, device="cuda") for k in self.COLUMNS)
            self._dev_dirty = self._live_ranges()
        for start, count in self._dev_dirty:
            for k, d in zip(self.COLUMNS, self._dev):
                host = np.ascontiguousarray(self._cols[k][start:start + count], dtype=np.float32)
                d[start:start + count].copy_(torch.from_numpy(host))
        self._dev_dirty = []
        return self._dev, self._capacity

    # ---- sampling ----
    def sample_indices(self, minibatch_size: int) -> np.ndarray:
        """The reference's draw (replay_buffer.py:58): logical indices, with replacement, global numpy RNG."""
        return np.random.randint(0, self.current_size, minibatch_size)

    def gather(self, indices: np.ndarray) -> Dict[str, np.ndarray]:
        """Rows at logical ``indices`` (any shape); the result has the shapes / dtypes ``sample_minibatch`` returns."""
        phys = self._physical(np.asarray(indices))
        return {k: self._cols[k][phys] for k in self.COLUMNS}

    def sample_minibatch(self, minibatch_size: int = 32) -> Dict[str, np.ndarray]:
        return self.gather(self.sample_indices(minibatch_size))

    # ---- reference-compatible read-only views (logical order) ----
    def _view(self, k: str) -> List:
        if self._cols[k] is None:
            return []
        return list(self._cols[k][self._physical(np.arange(self.current_size))])

    observations = property(lambda self: self._view("observations"))
    actions = property(lambda self: self._view("actions"))
    rewards = property(lambda self: [float(x) for x in self._view("rewards")])
    next_observations = property(lambda self: self._view("next_observations"))
    dones = property(lambda self: [bool(x) for x in self._view("dones")])
