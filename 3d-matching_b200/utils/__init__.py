"""Mirror of the reference's src/utils package: only what src/main.py imports (setup_logging).  The wall-clock
Profiler (src/utils/profiler.py) is replaced by CUDA-event statistics: pcr_kernel_stats and benchmark_ransac.EventProfiler."""
