"""`Logger` and `Losses`, the two names the reference's scripts import from the top-level `utils` module
(reference: utils.py:9-42).  Host-side conveniences, not on the integration path."""
import pathlib

import torch

try:
    from tqdm import tqdm
    _echo = tqdm.write
except Exception:   # pragma: no cover
    _echo = print


class Logger:
    """Two text logs (training / testing); every line is also echoed through tqdm.write."""

    def __init__(self, logfile_path: pathlib.Path, train_log_name=None, test_log_name=None):
        base = pathlib.Path(logfile_path)
        self.training_logfile = open(base / train_log_name, "w") if train_log_name is not None else None
        self.testing_logfile = open(base / test_log_name, "w") if test_log_name is not None else None

    def _emit(self, fh, strs):
        line = " ".join(strs)
        if fh is not None:
            fh.write(line + "\n")
        _echo(line)

    def training_log(self, *strs):
        self._emit(self.training_logfile, strs)

    def testing_log(self, *strs):
        self._emit(self.testing_logfile, strs)

    def close(self):
        for fh in (self.training_logfile, self.testing_logfile):
            if fh is not None and not fh.closed:
                fh.close()

    def __del__(self):
        self.close()


class Losses:
    """NaN / blow-up guard for a vector of per-series losses (reference: utils.py:29-42)."""

    def __init__(self, log: Logger):
        self.logger = log

    def multi_time_series_loss(self, loss: torch.Tensor, limit_loss=None):
        normalised = torch.where(loss < 1.0e-6, loss, loss / loss.detach())
        if torch.isnan(loss).any():
            self.logger.training_log(f"wrong loss: {loss.detach()}")
            return torch.sum(loss - loss)
        if limit_loss is not None and torch.any(loss > 1):
            if torch.any(loss > limit_loss):
                self.logger.training_log(f"too big loss: {loss.detach()}")
                return torch.sum(normalised)
            return torch.sum(loss)
        return torch.sum(normalised)
