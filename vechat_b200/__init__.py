"""vechat_b200: the POA window-correction path of VeChat on B200."""
import os

# The engine runs up to 48 stream groups side by side; with the default 8 hardware queues streams alias and serialise
# on each other's dependencies.  Must be set before the CUDA context exists (csrc/vgc_engine.cu ConnectionsInit does
# the same when libvgc.so is loaded first).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
