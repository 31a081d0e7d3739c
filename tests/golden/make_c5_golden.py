#!/usr/bin/env python
"""Golden for BASELINE config C5 (synthetic gray 3840x2160, D = 0..384) made with the oracle port
(oracle/liboracle.so, bit-identical to the unmodified reference on every stage, tests/test_oracle_golden.py).

    python tests/golden/make_c5_golden.py        # ~26 GB of host RAM, 10-20 minutes on 8 cores

The full float map would be 33 MB, so the fixture keeps every 4th row of the final map and of both WTA maps,
the full-resolution invalid-pixel mask (bit-packed) and SHA-256 digests of the full maps:
tests/golden/port_c5_gray4k_d384.npz.  tests/test_gpu_fullsize.py compares the CUDA path with it."""
import hashlib
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from tea_stereo_matching_b200.synth import synth_v1  # noqa: E402

H, W, D, SEED, STEP = 2160, 3840, 384, 3000, 4


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    left, right = synth_v1(H, W, D, seed=SEED, gray=True)
    t0 = time.time()
    st = oracle.Port().run(left, right, D, volumes=False)
    print(f"port: {time.time() - t0:.0f} s", st.seconds, flush=True)
    np.savez_compressed(
        Path(__file__).resolve().parent / "port_c5_gray4k_d384.npz",
        H=H, W=W, max_disparity=D, seed=SEED, row_step=STEP,
        final_rows=st.final[::STEP].copy(),
        wta0_rows=st.wta[0][::STEP].astype(np.int16), wta1_rows=st.wta[1][::STEP].astype(np.int16),
        invalid_bits=np.packbits(st.final < 0),
        final_sha256=np.array(sha(st.final)), wta0_sha256=np.array(sha(st.wta[0])), wta1_sha256=np.array(sha(st.wta[1])),
        left_sha256=np.array(sha(left)), right_sha256=np.array(sha(right)),
    )


if __name__ == "__main__":
    main()
