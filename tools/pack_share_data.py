#!/usr/bin/env python
"""Pack the reference's *data* tables that the hot path consumes into one .npz.

Run in the build container only (needs /root/reference):

    python tools/pack_share_data.py [/root/reference/share] [data/egg_share.npz]

Packs (values untouched, f64/f32 exactly as stored in the FITS files):
  * every filter curve of share/filter-db (db.dat order)  -> filter/<name>/lam, /res
  * share/ir_lib_ce01.fits                                -> ce01/lam, ce01/sed
  * share/mass_func_candels.fits                          -> mf/zb, mf/mb, mf/active, mf/passive
The GPU box has no /root/reference, so tests, smoke() and bench.py read this
pack instead.  No reference *source* is copied; these are input data tables.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from egg_b200.fitscol import read_filter_db, read_table  # noqa: E402


def main():
    share = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/share"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "data", "egg_share.npz")
    arrays = {}
    db = read_filter_db(os.path.join(share, "filter-db", "db.dat"))
    names = []
    for name, path in db.items():
        t = read_table(path)
        arrays[f"filter/{name}/lam"] = t["LAM"]
        arrays[f"filter/{name}/res"] = t["RES"]
        names.append(name)
    arrays["filter_names"] = np.array(names)
    t = read_table(os.path.join(share, "ir_lib_ce01.fits"))
    arrays["ce01/lam"] = t["LAM"]
    arrays["ce01/sed"] = t["SED"]
    t = read_table(os.path.join(share, "mass_func_candels.fits"))
    arrays["mf/zb"] = t["ZB"]
    arrays["mf/mb"] = t["MB"]
    arrays["mf/active"] = t["ACTIVE"]
    arrays["mf/passive"] = t["PASSIVE"]
    np.savez_compressed(out, **arrays)
    print(f"wrote {out}: {len(names)} filters, {os.path.getsize(out)/1e6:.2f} MB")


if __name__ == "__main__":
    main()
