"""Feature normalisation on the host side: the reference's ``reader.preprocess`` surface.

``cmn`` (reference reader/preprocess.py:34-41) and ``GlobalMeanVarianceNormalization``
(:89-229; pickled to exp_dir/transform.pkl at bin/train_ce.py:107-108 and unpickled at
bin/train_se.py:103-106 -- the attribute names below are the pickle contract).  On the GPU path
these two are fused into the feature gather (pk2_gather_norm); the numpy/torch versions here
serve host-side callers and statistics estimation.
"""
import numpy as np
import torch


def cmn(data, axis=1, is_tensor=False):
    """Subtract the mean along ``axis`` (per-utterance cepstral mean normalisation)."""
    if is_tensor:
        return data - data.mean(dim=axis, keepdim=True)
    return data - data.mean(axis=axis, keepdims=True)


class GlobalMeanVarianceNormalization(object):
    """Global mean / variance normalisation with statistics learned from a sample of the data."""

    STD_FLOOR = 1e-2

    def __init__(self, mean_vec=None, std_vec=None, mean_norm=True, var_norm=True):
        self.mean_vec = mean_vec
        self.std_vec = std_vec
        self.mean_vec_tensor = None
        self.std_vec_tensor = None
        self.mean_norm = mean_norm
        self.var_norm = var_norm
        self.mean_stats = None
        self.var_stats = None
        self.n_frame = 0

    # -- statistics ------------------------------------------------------------------
    def initialize_stats(self, dim):
        self.mean_stats = np.zeros((dim, 1), np.float32)
        self.var_stats = np.zeros((dim, 1), np.float32)
        self.n_frame = 0

    def accumulate_stats(self, data):
        """data: [T, D]"""
        if self.mean_stats is None:
            self.initialize_stats(data.shape[1])
        self.mean_stats += data.sum(axis=0)[:, None]
        self.var_stats += np.square(data).sum(axis=0)[:, None]
        self.n_frame += data.shape[0]

    def learn_mean_and_variance_from_stats(self):
        mean = self.mean_stats / self.n_frame
        std = np.sqrt(self.var_stats / self.n_frame - np.square(mean))
        std = np.maximum(std, self.STD_FLOOR)
        std[~np.isfinite(std)] = 1.0
        self.mean_vec = mean.astype(np.float32).T          # row vectors [1, D]
        self.std_vec = std.astype(np.float32).T

    def learn_mean_and_variance_from_train_loader(self, train_set, stream_keys=(), n_sample_to_use=200):
        n = len(train_set)
        picks = np.arange(n) if n <= n_sample_to_use else np.random.choice(n, n_sample_to_use, replace=False)
        used = 0
        for idx in picks:
            item = train_set[int(idx)]
            for key in stream_keys:
                streams = item[key] if isinstance(item[key], list) else [item[key]]
                for mat in streams:
                    if isinstance(mat, np.ndarray):
                        self.accumulate_stats(mat)
                        used += 1
            if used > n_sample_to_use:
                break
        self.learn_mean_and_variance_from_stats()

    # -- application -----------------------------------------------------------------
    def apply_on_ndarray(self, data, is_tensor=False):
        out = data
        if self.mean_norm:
            if is_tensor:
                if self.mean_vec_tensor is None:
                    self.mean_vec_tensor = torch.tensor(self.mean_vec)
                out = out - self.mean_vec_tensor
            else:
                out = out - self.mean_vec
        if self.var_norm:
            if is_tensor:
                if self.std_vec_tensor is None:
                    self.std_vec_tensor = torch.tensor(self.std_vec)
                out = out / self.std_vec_tensor
            else:
                out = out / self.std_vec
        return out

    def device_vectors(self, device):
        """(mean [D], 1/std [D]) float32 tensors for the fused GPU gather (pk2_gather_norm)."""
        mean = torch.as_tensor(np.asarray(self.mean_vec, np.float32).reshape(-1), device=device)
        istd = torch.as_tensor((1.0 / np.asarray(self.std_vec, np.float32)).reshape(-1), device=device)
        if not self.mean_norm:
            mean = torch.zeros_like(mean)
        if not self.var_norm:
            istd = torch.ones_like(istd)
        return mean, istd
