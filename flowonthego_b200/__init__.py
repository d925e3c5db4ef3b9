"""flowonthego_b200 -- B200-native Dense Inverse Search optical flow (drop-in for the hot path of
zhaorz/FlowOnTheGo's CPU reference `kroeger/`).  See DESIGN.md and include/dis_c.h."""
import os as _os

# see ConnectionsDefault in csrc/engine.cu: must be in the environment before the process creates its CUDA
# context (e.g. before torch touches the GPU), hence at import time as well as at library load
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .api import (DisError, Engine, FlowStream, flow_epe, flow_to_color, write_png_bgr, OFClass, Params, PARAM_NAMES, lib, padded_size, pinned_empty, read_flo, read_image_bgr, read_image_gray,
                  run_dense, write_flo)

__all__ = ["DisError", "Engine", "OFClass", "Params", "PARAM_NAMES", "lib", "padded_size", "pinned_empty",
           "read_flo", "read_image_bgr", "read_image_gray", "run_dense", "write_flo", "flow_epe", "flow_to_color", "write_png_bgr", "FlowStream"]
