"""Scalar logging for the host-side learn() loops: one front end, pluggable sinks.

Same surface as the reference's manager (ref: metrics_manager.py:11-42: ``record_scalar(tag, scalar, total_steps,
tensorboard)``, ``dump()``, ``close()``, one ``tag: value`` line per scalar on stdout, an event file under
``<log_dir>/tensorboard``), but TensorBoard is optional here: the image that runs the GPU tests may not ship it."""
import logging
import os
import sys
from typing import List, Optional

logger = logging.getLogger(__name__)


class _StdoutSink:
    wants_all = True  # every scalar, with or without a step

    def write(self, tag: str, value: float, step: Optional[int]) -> None:
        sys.stdout.write("{}: {:<8.3g}\n".format(tag, value))

    def flush(self) -> None:
        sys.stdout.flush()

    def close(self) -> None:
        pass


class _TensorBoardSink:
    wants_all = False  # only scalars recorded with tensorboard=True

    def __init__(self, directory: str) -> None:
        from torch.utils.tensorboard import SummaryWriter  # ImportError -> the manager runs without this sink
        self.writer = SummaryWriter(directory)

    def write(self, tag: str, value: float, step: Optional[int]) -> None:
        if step is None:
            logger.warning("total_steps argument is required for tensorboard")
        self.writer.add_scalar(tag, value, step)

    def flush(self) -> None:
        self.writer.flush()

    def close(self) -> None:
        self.writer.close()


class MetricsManager:
    def __init__(self, log_dir: str = "."):
        self.log_dir = log_dir
        self._sinks: List[object] = [_StdoutSink()]
        self.tensorboard_writer = None  # the reference's attribute name; None when TensorBoard is unavailable
        try:
            board = _TensorBoardSink(os.path.join(log_dir, "tensorboard"))
        except Exception:
            board = None
        if board is not None:
            self._sinks.append(board)
            self.tensorboard_writer = board.writer

    def record_scalar(self, tag: str, scalar: float, total_steps: Optional[int] = None,
                      tensorboard: bool = False) -> None:
        for sink in self._sinks:
            if sink.wants_all or tensorboard:
                sink.write(tag, scalar, total_steps)

    def dump(self) -> None:
        for sink in self._sinks:
            sink.flush()

    def close(self) -> None:
        for sink in self._sinks:
            sink.close()
