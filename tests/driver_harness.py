"""Runs one of the reference's experiment drivers UNCHANGED (runpy on the file where it lies) against this repo's
`models` package, with stand-ins for the packages the image lacks (tests/driver_stubs: matplotlib, tensorboardX) and
for the dataset files (datasets).  Drivers without an epoch flag (ToyExperiments.py loops 10000 epochs) are stopped
through the clock they read: `timeit.default_timer` raises after DRIVER_MAX_TIMER_CALLS calls.

    python tests/driver_harness.py <reference root> <driver.py> [driver args...]
"""
import os
import runpy
import sys
import timeit

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


class _Stop(Exception):
    pass


def main():
    ref, driver = sys.argv[1], sys.argv[2]
    # our `models` first, then the stand-ins (they shadow the reference's `datasets`), then the reference's own `lib`
    sys.path[:0] = [REPO, os.path.join(HERE, "driver_stubs"), ref]
    limit = int(os.environ.get("DRIVER_MAX_TIMER_CALLS", "0"))
    if limit:
        real, calls = timeit.default_timer, [0]

        def limited():
            calls[0] += 1
            if calls[0] > limit:
                raise _Stop()
            return real()
        timeit.default_timer = limited
    sys.argv = [os.path.join(ref, driver)] + sys.argv[3:]
    try:
        runpy.run_path(os.path.join(ref, driver), run_name="__main__")
    except _Stop:
        print("DRIVER_STOPPED_BY_HARNESS", flush=True)
    import models
    import umnn_b200
    assert models.UMNNMAFFlow is umnn_b200.UMNNMAFFlow, "the driver did not run against this repository's package"
    print("DRIVER_OK", flush=True)


if __name__ == "__main__":
    main()
