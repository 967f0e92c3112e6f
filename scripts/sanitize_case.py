"""One small fused launch per kernel variant, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_fused_linear import run_case
for (m, k, n) in [(300, 768, 768), (300, 3072, 768), (200, 256, 96)]:
    run_case(m, k, n, 6, 6, True, 11)
    print("ok", m, k, n, flush=True)
