#!/usr/bin/env python3
"""Run named parity checks on the GPU box, one per process (a faulting kernel poisons the CUDA
context).  Usage: python tools/gpu_check.py <name>|--list ; results are printed as JSON lines."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import kernel_checks as kc
    from util import CASE_CFG
    names = list(kc.ALL) + ["model:" + c for c in CASE_CFG]
    if sys.argv[1:] == ["--list"]:
        print(" ".join(names))
        return 0
    rc = 0
    for name in sys.argv[1:]:
        t0 = time.time()
        try:
            if name.startswith("model:"):
                import model_checks as mc
                res = mc.run_case(name[6:])
                try:
                    mc.assert_case(name, res)
                    ok = True
                except AssertionError:
                    ok = False
            else:
                res = kc.ALL[name]()
                try:
                    kc.assert_ok(name, res)
                    ok = True
                except AssertionError:
                    ok = False
            print(json.dumps(dict(check=name, ok=ok, sec=round(time.time() - t0, 2), **res)), flush=True)
            rc |= 0 if ok else 1
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(check=name, ok=False, error=repr(e)[:400])), flush=True)
            traceback.print_exc()
            rc |= 2
    return rc


if __name__ == "__main__":
    sys.exit(main())
