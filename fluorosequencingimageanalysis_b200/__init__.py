"""B200-native spot hot path of FluorosequencingImageAnalysis (detection -> 2-D Gaussian PSF fits -> metrics ->
consolidation) behind the reference's own entry points.  See DESIGN.md / INTEGRATION.md."""
import os as _os

# The production pipeline (engine.FieldStream) keeps 5-8 batches in flight on their own CUDA streams.  CUDA maps streams
# onto CUDA_DEVICE_MAX_CONNECTIONS hardware work queues (default 8); with more streams than queues a host->device frame
# copy shares a queue with another batch's kernels and holds them back.  Measured on B200 (bench.py e2e region):
# 2.06e8 fits/s with the default, 2.31e8 -- equal to the device-resident rate -- with 32 queues.  The variable is read
# when the CUDA context is created, so it must be set before the first CUDA call of the process; an explicit setting
# by the user wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
