"""upfirdn2d achieved HBM GB/s on the SURVEY 8(d) inputs (bench.py's block) -- run under SG2_UPF_COLS_NHWC / SG2_UPF_COLS for sweeps."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench     # noqa: E402

print(os.environ.get('SG2_UPF_COLS_NHWC', '-'), os.environ.get('SG2_UPF_COLS', '-'),
      json.dumps({k.split(' [')[0][:40]: (v['frac'] if isinstance(v, dict) else v) for k, v in bench.upfirdn2d_rates(torch.device('cuda'), bench.load_peaks()).items()}))
