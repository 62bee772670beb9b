"""Generate the committed fixtures under tests/golden/ from the read-only reference tree.

Runs ONLY in the build container (needs /root/reference).  Nothing at test / bench
run time reads /root/reference: the outputs of this script are committed.

What it writes
--------------
cance_inputs.npz   Cance model inputs rebuilt the way the reference builds them
                   (smash/core/_read_input_data.py:150-343, _build_model.py:233-257,
                   tools/raster_handler.py:gdal_read_windowed_raster) without GDAL/h5py:
                   mesh arrays, prcp (T,nrow,ncol) f32, daily PET maps + per-step
                   (day, ratio) so pet(t) = f32(pet_daily[day[t]] * ratio[t]), qobs, descriptors.
cance_golden.npz   the hot-path keys of smash/tests/baseline.hdf5 (run.cost, multiple_run.*,
                   mutiple_run.slc_*, optimize.* costs / maps) + the sample matrix of
                   generate_samples(problem, n=10, random_state=99)
                   (smash/core/generate_samples.py:357-364).
france_mesh.npz    mesh_France.hdf5 (flwdir, flwacc, active_cell, path as stored) + the
                   golden bbox_mesh.{flwdir,flwacc} / xy_mesh.* integer contract.

Usage:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import glob
import os
import struct
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _h5lite import H5File  # noqa: E402

REF = "/root/reference/smash"

# smash/core/_constant.py:49-77 (float32 table)
RATIO_PET_HOURLY = np.array(
    [0, 0, 0, 0, 0, 0, 0, 0.035, 0.062, 0.079, 0.097, 0.11, 0.117, 0.117, 0.11, 0.097,
     0.079, 0.062, 0.035, 0, 0, 0, 0, 0], dtype=np.float32)


# ----------------------------------------------------------------------------- GeoTIFF
def _tiff_geo(path):
    """(xleft, ytop, xres, yres) of the upper-left pixel CORNER, GDAL semantics
    (PixelIsPoint rasters are shifted by half a pixel)."""
    b = open(path, "rb").read()
    bo = "<" if b[:2] == b"II" else ">"
    (off,) = struct.unpack(bo + "I", b[4:8])
    (n,) = struct.unpack(bo + "H", b[off:off + 2])
    tags = {}
    for i in range(n):
        tag, typ, cnt, val = struct.unpack(bo + "HHII", b[off + 2 + 12 * i:off + 14 + 12 * i])
        tags[tag] = (typ, cnt, val, off + 2 + 12 * i + 8)
    def doubles(tag):
        typ, cnt, val, _ = tags[tag]
        return struct.unpack(bo + "d" * cnt, b[val:val + 8 * cnt])
    def shorts(tag):
        typ, cnt, val, inl = tags[tag]
        o = val if cnt * 2 > 4 else inl
        return struct.unpack(bo + "H" * cnt, b[o:o + 2 * cnt])
    sx, sy, _ = doubles(33550)
    tp = doubles(33922)
    xleft = tp[3] - tp[0] * sx
    ytop = tp[4] + tp[1] * sy
    if 34735 in tags:
        gk = shorts(34735)
        for k in range(4, len(gk), 4):
            if gk[k] == 1025 and gk[k + 3] == 2:  # GTRasterTypeGeoKey == RasterPixelIsPoint
                xleft -= 0.5 * sx
                ytop += 0.5 * sy
    nodata = None
    if 42113 in tags:
        typ, cnt, val, inl = tags[42113]
        o = val if cnt > 4 else inl
        try:
            nodata = float(b[o:o + cnt].split(b"\0")[0].decode())
        except ValueError:
            nodata = None
    return xleft, ytop, sx, sy, nodata


def read_windowed(path, mesh, lacuna=-99.0):
    """Restates gdal_read_windowed_raster for rasters already at the mesh resolution."""
    import cv2

    xleft, ytop, sx, sy, nodata = _tiff_geo(path)
    assert sx == mesh["dx"] and sy == mesh["dx"], "reprojection path not needed for the shipped data"
    col_off = (mesh["xmin"] - xleft) / sx
    row_off = (ytop - mesh["ymax"]) / sy
    # GDAL ReadAsArray with float offsets: int(off + 0.5)
    c0 = int(col_off + 0.5)
    r0 = int(row_off + 0.5)
    a = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    win = a[r0:r0 + mesh["nrow"], c0:c0 + mesh["ncol"]]
    out = win.astype("float64")
    if nodata is not None:
        out[win == nodata] = lacuna
    return out


# ----------------------------------------------------------------------------- Cance inputs
def build_cance():
    ds = f"{REF}/dataset/Cance"
    f = H5File(f"{ds}/mesh_Cance.hdf5")
    mesh = dict(f.attrs)
    mesh.pop("_save_func", None)
    for k in f.keys():
        mesh[k] = f[k]
    start, end, dt = "2014-09-15 00:00", "2014-11-14 00:00", 3600
    date_range = pd.date_range(start=start, end=end, freq=f"{dt}s")[1:]
    T = len(date_range)
    nrow, ncol = mesh["nrow"], mesh["ncol"]

    # prcp (_read_input_data.py:150-200)
    files = sorted(glob.glob(f"{ds}/prcp/**/*tif*", recursive=True))
    prcp = np.full((T, nrow, ncol), -99.0, np.float32)
    for i, date in enumerate(date_range):
        key = date.strftime("%Y%m%d%H%M")
        hit = [p for p in files if key in p]
        if hit:
            prcp[i] = (read_windowed(hit[0], mesh) * 0.1).astype(np.float32)
            files.remove(hit[0])

    # daily inter-annual PET (_read_input_data.py:203-276): leap-year day-of-year match
    pfiles = sorted(glob.glob(f"{ds}/pet/**/*tif*", recursive=True))
    leap_days = pd.date_range(start="202001010000", end="202012310000", freq="1D")
    doy = date_range.day_of_year
    pet_daily, pet_day, pet_ratio = [], np.full(T, -1, np.int32), np.zeros(T, np.float32)
    for day in leap_days:
        if day.day_of_year in doy:
            hit = [p for p in pfiles if day.strftime("%m%d") in p]
            ind_day = np.where(day.day_of_year == doy)[0]
            assert hit, f"missing PET file {day}"
            pet_daily.append(read_windowed(hit[0], mesh) * 1)
            sub = date_range[ind_day]
            for j in range(24):
                step = day + pd.Timedelta(seconds=j * dt)
                ind_step = sub.indexer_at_time(step)
                pet_day[ind_day[ind_step]] = len(pet_daily) - 1
                pet_ratio[ind_day[ind_step]] = RATIO_PET_HOURLY[j]
            pfiles.remove(hit[0])
    assert (pet_day >= 0).all()
    pet_daily = np.stack(pet_daily)

    # qobs (_read_input_data.py:20-75)
    st = pd.Timestamp(start)
    qobs = np.full((mesh["ng"], T), -99.0, np.float32)
    for i, code in enumerate(mesh["code"].astype("U")):
        path = glob.glob(f"{ds}/qobs/**/*{code}*.csv", recursive=True)[0]
        with open(path) as fh:
            header = pd.Timestamp(fh.readline())
            skip = int((st - header).total_seconds() / dt) + 1
            k = 0
            for j, line in enumerate(fh):
                if j >= skip:
                    if k >= T:
                        break
                    qobs[i, k] = float(line)
                    k += 1

    desc = np.zeros((nrow, ncol, 2), np.float32)
    for i, name in enumerate(["slope", "dd"]):
        desc[..., i] = read_windowed(f"{ds}/descriptor/{name}.tif", mesh).astype(np.float32)

    np.savez_compressed(
        f"{HERE}/cance_inputs.npz",
        dx=np.float32(mesh["dx"]), dt=np.float32(dt), nrow=nrow, ncol=ncol, ng=mesh["ng"], nac=mesh["nac"],
        flwdir=mesh["flwdir"].astype(np.int32), flwacc=mesh["flwacc"].astype(np.int32),
        active_cell=mesh["active_cell"].astype(np.int32), path=mesh["path"].astype(np.int32),
        gauge_pos=mesh["gauge_pos"].astype(np.int32), area=mesh["area"].astype(np.float32),
        code=mesh["code"], flwdst=mesh["flwdst"].astype(np.float32),
        prcp=prcp, pet_daily=pet_daily, pet_day=pet_day, pet_ratio=pet_ratio,
        qobs=qobs, descriptor=desc,
    )
    print("cance_inputs.npz: T =", T, "prcp max", prcp.max(), "pet days", len(pet_daily))


# ----------------------------------------------------------------------------- golden outputs
def build_golden():
    g = H5File(f"{REF}/tests/baseline.hdf5")
    keep = [k for k in g.keys() if k.split(".")[0] in ("run", "multiple_run", "mutiple_run", "optimize", "bayes_estimate", "bayes_optimize", "ann_optimize_1", "ann_optimize_2", "gen_samples", "net_init")
            or k.startswith("xy_mesh.") or k.startswith("mesh_io.")]
    out = {k: g[k] for k in keep}
    # generate_samples(problem, n=10, random_state=99): one legacy-uniform draw per variable
    bounds = [(1e-6, 1e3), (1e-6, 1e3), (-50.0, 50.0), (1e-6, 1e3)]  # cp, cft, exc, lr
    np.random.seed(99)
    out["samples.cp_cft_exc_lr"] = np.stack([np.random.uniform(lo, hi, 10) for lo, hi in bounds])
    np.savez_compressed(f"{HERE}/cance_golden.npz", **out)
    print("cance_golden.npz:", len(out), "keys")

    f = H5File(f"{REF}/dataset/France/mesh_France.hdf5")
    path = f["path"]
    np.savez_compressed(
        f"{HERE}/france_mesh.npz",
        dx=np.float32(f.attrs["dx"]), nrow=f.attrs["nrow"], ncol=f.attrs["ncol"], nac=f.attrs["nac"],
        flwdir=f["flwdir"].astype(np.int16), flwacc=f["flwacc"].astype(np.int32),
        active_cell=f["active_cell"].astype(np.int8), path=path.astype(np.int16),
    )
    # the golden bbox_mesh.{flwdir,flwacc} of baseline.hdf5 (smash/tests/mesh/test_meshing.py:44-60) are
    # bit-identical to the shipped mesh_France.hdf5 arrays, so they are not stored twice.
    assert (g["bbox_mesh.flwdir"] == f["flwdir"]).all() and (g["bbox_mesh.flwacc"] == f["flwacc"]).all()
    print("france_mesh.npz written")


if __name__ == "__main__":
    build_cance()
    build_golden()
