"""Helper of test_oracle_cpu.py::test_every_entry_point_survives_null_arguments (run in a subprocess: a missing check would be a
segmentation fault, not an exception).  Calls every exported function with NULL for every pointer and 0 for every scalar and
prints one line per entry point: name, return code."""
import ctypes as C
import sys

from malio_b200 import capi

lib = capi.load()
for name in capi.EXPORTS:
    fn = getattr(lib, name)
    at = fn.argtypes
    if at is None:
        continue
    args = []
    for t in at:
        if t in (C.c_float, C.c_double):
            args.append(0.0)
        elif t in (C.c_int, C.c_uint32, C.c_uint64, C.c_int32, C.c_int64, C.c_uint8, C.c_uint16):
            args.append(0)
        else:
            args.append(None)
    print(name, end=" ", flush=True)
    rc = fn(*args)
    print(rc if isinstance(rc, int) else "ptr", flush=True)
print("DONE")
