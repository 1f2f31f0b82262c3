#!/bin/bash
# (kept as the entry point the queued gpurun call names) -> round-2 session A
exec bash tools/gpu_r2_a.sh
