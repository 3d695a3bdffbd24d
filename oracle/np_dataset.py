"""numpy restatement of the City-dataset .bin layouts and the two preprocess handlers.  TEST INFRASTRUCTURE ONLY.

  file_player/src/ROSThread.cpp:776-795 (Livox: 17-byte records), :952-967 (Ouster: 22-byte records)
  MA_LIO/src/preprocess.cpp:59-110 (avia_handler), :112-152 (oust64_handler)
Parity unpinned by the reference (no fixtures); the formats are read off the player's source."""
import numpy as np

LIVOX_REC = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("reflectivity", "u1"), ("tag", "u1"), ("line", "u1"), ("t16", "<u2")])
OUSTER_REC = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("ring", "<u2"), ("t", "<u4")])
assert LIVOX_REC.itemsize == 17 and OUSTER_REC.itemsize == 22


def read_records(path, rec, eof_quirk=True):
    raw = np.fromfile(path, dtype=np.uint8)
    n = raw.size // rec.itemsize
    out = raw[: n * rec.itemsize].view(rec)
    if eof_quirk:
        out = np.concatenate([out, np.zeros(1, rec)])
    return out


def avia_handler(p, n_scans=6, point_filter_num=1, blind=0.5):
    out, inten = [], []
    valid = 0
    full = np.zeros((len(p), 3), np.float32)
    for i in range(1, len(p)):
        if p["line"][i] < n_scans and ((p["tag"][i] & 0x30) == 0x10 or (p["tag"][i] & 0x30) == 0x00):
            valid += 1
            if valid % point_filter_num == 0:
                full[i] = (p["x"][i], p["y"][i], p["z"][i])
                curv = np.float32(p["t16"][i]) / np.float32(1000000)
                if curv > 100:
                    continue
                d = np.abs(full[i] - full[i - 1])
                rng = np.float64(full[i, 0] * full[i, 0] + full[i, 1] * full[i, 1] + full[i, 2] * full[i, 2])
                if d[0] > 1e-7 or d[1] > 1e-7 or (d[2] > 1e-7 and rng > blind * blind):
                    out.append((full[i, 0], full[i, 1], full[i, 2], curv))
                    inten.append(np.float32(p["reflectivity"][i]))
    return np.array(out, np.float32).reshape(-1, 4), np.array(inten, np.float32)


def oust64_handler(p, point_filter_num=1, blind=0.5, time_unit_scale=1e-3):
    out, inten = [], []
    for i in range(len(p)):
        if i % point_filter_num != 0:
            continue
        x, y, z = p["x"][i], p["y"][i], p["z"][i]
        if np.float64(x * x + y * y + z * z) < blind * blind:
            continue
        out.append((x, y, z, np.float32(p["t"][i]) * np.float32(time_unit_scale) * np.float32(1e-9)))
        inten.append(p["intensity"][i])
    return np.array(out, np.float32).reshape(-1, 4), np.array(inten, np.float32)
