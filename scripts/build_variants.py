"""Build experiment variants of the library (compile-time switches of the tensor-core kernels) under
umnn_b200/variants/; select one at run time with UMNN_B200_LIB=<path>."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from umnn_b200 import build as B   # noqa: E402

VARIANTS = {
    "lds_count": ["-DUMNN_TC_WAIT_STYLE=0"],                                 # LDS, hint, clock every 4096 polls
    "gen_count": ["-DUMNN_TC_SMEM_GENERIC=1", "-DUMNN_TC_WAIT_STYLE=0"],
    "lds_timer": ["-DUMNN_TC_WAIT_STYLE=1"],                                 # = product build
    "gen_timer": ["-DUMNN_TC_SMEM_GENERIC=1", "-DUMNN_TC_WAIT_STYLE=1"],    # = visit r1i
    "lds_plain": ["-DUMNN_TC_WAIT_STYLE=2", "-DUMNN_TC_WAIT_HINT=0"],
    "gen_plain": ["-DUMNN_TC_SMEM_GENERIC=1", "-DUMNN_TC_WAIT_STYLE=2", "-DUMNN_TC_WAIT_HINT=0"],   # = before this change
    "lds_plain_hint": ["-DUMNN_TC_WAIT_STYLE=2"],
    "f32x2_on": [],                                                          # = product build
    "f32x2_off": ["-DUMNN_TC_F32X2=0"],
    "lds_sleep20": ["-DUMNN_TC_WAIT_STYLE=3", "-DUMNN_TC_WAIT_SLEEP_NS=20"],
    "lds_sleep50": ["-DUMNN_TC_WAIT_STYLE=3", "-DUMNN_TC_WAIT_SLEEP_NS=50"],
    "lds_sleep100": ["-DUMNN_TC_WAIT_STYLE=3", "-DUMNN_TC_WAIT_SLEEP_NS=100"],
    "lds_sleep200": ["-DUMNN_TC_WAIT_STYLE=3", "-DUMNN_TC_WAIT_SLEEP_NS=200"],
    "issue_workconserving": ["-DUMNN_TC_ISSUE_WORKCONSERVING=1"],            # segment 1 advances while segment 0 waits (no gain measured: visit r2l)
    "lds_sleep50_nohint": ["-DUMNN_TC_WAIT_STYLE=3", "-DUMNN_TC_WAIT_SLEEP_NS=50", "-DUMNN_TC_WAIT_HINT=0"],
}

if __name__ == "__main__":
    out_dir = os.path.join(REPO, "umnn_b200", "variants")
    os.makedirs(out_dir, exist_ok=True)
    for name in (sys.argv[1:] or VARIANTS):
        path = os.path.join(out_dir, f"libumnn_b200_{name}.so")
        B.build(force=True, extra_flags=VARIANTS[name], out=path)
        print(path, flush=True)
