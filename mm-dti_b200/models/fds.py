"""Drop-in for the reference's models/fds.py: ``FDS`` with the same constructor, buffers
(state_dict names), ``smooth`` / ``update_last_epoch_stats`` / ``update_running_stats`` /
``reset``, running on the K5 kernels (ops_fds).  Differences are in HOW only:

  * the reference bins each sample on the host (one device->host sync per molecule,
    models/fds.py:125,164) and loops over torch.unique(bins); here binning, grouping, the
    per-bucket mean / unbiased variance, the EMA and the calibration stay on the device;
  * ``device`` is honoured (the reference hard-codes .to('cuda') for the window, :84);
  * ``dp`` (dist.DataParallelCtx): the epoch statistics are all-reduced so every rank keeps
    identical buffers, and ``smooth`` sees the bins of the global batch.
Aliasing of running_* and running_*_last_epoch after the first update (fds.py:87-88) is kept."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops_fds


def anomaly_clean_regression(data):
    """3-sigma cleaning of the (scaled) regression targets (fds.py:18-29)."""
    mean, std = data.mean(), data.std()
    return data[(data > mean - 3 * std) & (data < mean + 3 * std)]


def _gaussian_taps(sigma, radius):
    xs = np.arange(-radius, radius + 1, dtype=np.float64)
    w = np.exp(-0.5 * (xs / float(sigma)) ** 2)
    return w / w.sum()


def kernel_window(kernel, ks, sigma):
    """FDS._get_kernel_window (fds.py:69-84) as a float32 numpy array.  'gaussian' is
    scipy.ndimage.gaussian_filter1d of a unit impulse of length ks (truncate 4, boundary 'reflect'
    = half-sample symmetric), normalised to sum 1."""
    assert kernel in ['gaussian', 'triang', 'laplace']
    half = (ks - 1) // 2
    if kernel == 'gaussian':
        radius = int(4.0 * float(sigma) + 0.5)
        taps = _gaussian_taps(sigma, radius)
        impulse = np.zeros(ks, dtype=np.float32)
        impulse[half] = 1.0
        pos = np.arange(-radius, ks + radius) % (2 * ks)
        pos = np.where(pos >= ks, 2 * ks - 1 - pos, pos)
        ext = impulse.astype(np.float64)[pos]
        resp = np.array([np.dot(ext[i:i + 2 * radius + 1], taps) for i in range(ks)])
        win = resp / resp.sum()
    elif kernel == 'triang':
        n = np.arange(1, ks + 1, dtype=np.float64)
        tri = 1 - np.abs(n - (ks + 1) / 2) / ((ks + 1) / 2) if ks % 2 else 1 - np.abs(n - (ks + 1) / 2) / (ks / 2)
        win = tri / tri.sum()
    else:
        lap = np.exp(-np.abs(np.arange(-half, half + 1, dtype=np.float64)) / sigma) / (2. * sigma)
        win = lap / lap.sum()
    return win.astype(np.float32)


class FDS(nn.Module):

    def __init__(self, feature_dim, raw_data, col_data, using_scale, bucket_num=100, bucket_start=0, start_update=0,
                 start_smooth=1, kernel='gaussian', ks=5, sigma=2, momentum=0.9, device='cuda'):
        super(FDS, self).__init__()
        self.feature_dim = feature_dim
        self.bucket_num = bucket_num
        self.bucket_start = bucket_start
        self.half_ks = (ks - 1) // 2
        self.momentum = momentum
        self.start_update = start_update
        self.start_smooth = start_smooth
        self.device = device
        self.dp = None
        if isinstance(raw_data, str):
            import pandas as pd
            self.raw_data = pd.read_csv(raw_data).loc[:, col_data].values
        else:                           # an array of training targets may be passed directly
            self.raw_data = np.asarray(raw_data)
        values = np.array(self.raw_data, dtype=np.float64, copy=True).reshape(-1)
        if using_scale:                 # StandardScaler (population std) then the 3-sigma clean
            std = values.std()
            values = (values - values.mean()) / (std if std > 0 else 1.0)
            values = anomaly_clean_regression(values)
        self.min_value = np.min(values)
        self.bin_width = (np.max(values) - np.min(values)) / bucket_num

        nb = bucket_num - bucket_start
        self.register_buffer('epoch', torch.zeros(1).fill_(start_update))
        self.register_buffer('running_mean', torch.zeros(nb, feature_dim))
        self.register_buffer('running_var', torch.ones(nb, feature_dim))
        self.register_buffer('running_mean_last_epoch', torch.zeros(nb, feature_dim))
        self.register_buffer('running_var_last_epoch', torch.ones(nb, feature_dim))
        self.register_buffer('smoothed_mean_last_epoch', torch.zeros(nb, feature_dim))
        self.register_buffer('smoothed_var_last_epoch', torch.ones(nb, feature_dim))
        self.register_buffer('num_samples_tracked', torch.zeros(nb))
        self.kernel_window = self._get_kernel_window(kernel, ks, sigma)
        self._epoch_host = float(start_update)      # host mirror of `epoch`: no device sync in the hot path

    @staticmethod
    def _get_kernel_window(kernel, ks, sigma):
        return torch.from_numpy(kernel_window(kernel, ks, sigma))

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.kernel_window = fn(self.kernel_window)
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._epoch_host = float(self.epoch.item())
        return out

    def _update_last_epoch_stats(self):
        self.running_mean_last_epoch = self.running_mean            # aliases, like the reference
        self.running_var_last_epoch = self.running_var
        win = self.kernel_window.to(self.running_mean.device)
        self.smoothed_mean_last_epoch = ops_fds.fds_window_smooth(self.running_mean_last_epoch, win)
        self.smoothed_var_last_epoch = ops_fds.fds_window_smooth(self.running_var_last_epoch, win)

    def reset(self):
        self.running_mean.zero_()
        self.running_var.fill_(1)
        self.running_mean_last_epoch.zero_()
        self.running_var_last_epoch.fill_(1)
        self.smoothed_mean_last_epoch.zero_()
        self.smoothed_var_last_epoch.fill_(1)
        self.num_samples_tracked.zero_()

    def update_last_epoch_stats(self, epoch):
        if epoch == self._epoch_host + 1:
            self.epoch += 1
            self._epoch_host += 1
            self._update_last_epoch_stats()

    def _bins(self, labels):
        return ops_fds.fds_bin(labels, float(self.min_value), float(self.bin_width), self.bucket_start, self.bucket_num,
                               dp=self.dp)

    def update_running_stats(self, features, labels, epoch):
        if epoch < self._epoch_host:
            return
        l0 = ops_fds.label_column(labels)
        assert self.feature_dim == features.size(1), "Input feature dimension is not aligned!"
        assert features.size(0) == l0.size(0), "Dimensions of features and labels are not aligned!"
        l0 = l0.to(features.device)
        bins, present = self._bins(l0)
        ops_fds.fds_update_running_stats(features, bins, present, self.bucket_start, self.bucket_num, self.running_mean,
                                         self.running_var, self.num_samples_tracked, self.momentum,
                                         epoch == self.start_update, dp=self.dp)

    def smooth(self, features, labels, epoch):
        if epoch < self.start_smooth:
            return features
        bins, present = self._bins(labels.to(features.device))
        return ops_fds.FDSSmoothFn.apply(features, bins, present, self.running_mean_last_epoch, self.running_var_last_epoch,
                                         self.smoothed_mean_last_epoch, self.smoothed_var_last_epoch, self.bucket_start,
                                         self.bucket_num)
