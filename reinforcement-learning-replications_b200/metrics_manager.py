import logging
import os
import sys
from typing import Optional

logger = logging.getLogger(__name__)


class MetricsManager:
    """stdout + TensorBoard scalar sink with the reference's interface (ref: metrics_manager.py:11-42)."""

    def __init__(self, log_dir: str = "."):
        self.log_dir = log_dir
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.tensorboard_writer = SummaryWriter(os.path.join(log_dir, "tensorboard"))
        except Exception:  # tensorboard is optional in this image
            self.tensorboard_writer = None

    def record_scalar(self, tag: str, scalar: float, total_steps: Optional[int] = None, tensorboard: bool = False) -> None:
        print("{}: {:<8.3g}".format(tag, scalar))
        if tensorboard and self.tensorboard_writer is not None:
            if total_steps is None:
                logger.warning("total_steps argument is required for tensorboard")
            self.tensorboard_writer.add_scalar(tag, scalar, total_steps)

    def dump(self) -> None:
        sys.stdout.flush()
        if self.tensorboard_writer is not None:
            self.tensorboard_writer.flush()

    def close(self) -> None:
        if self.tensorboard_writer is not None:
            self.tensorboard_writer.close()
