"""One tensor-core conv launch for ncu: python profiles/profile_tc_one.py C T k mode [f8]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from promonet_b200.tc_probe import run_tc_conv  # noqa: E402

channels, t_len, kernel = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = sys.argv[4] if len(sys.argv) > 4 else 'c2'
f8 = len(sys.argv) > 5 and sys.argv[5] == 'f8'
print(run_tc_conv(32, channels, t_len, kernel, mode, repeats=2, f8=f8))
