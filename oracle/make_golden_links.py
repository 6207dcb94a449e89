"""Generate tests/golden/link_programs.npz (build container only; TEST INFRASTRUCTURE ONLY).

Every film / conductance closure of the reference's library (thermca/lib/{free_conv,forced_conv,radiation,
combd_film,evaporation}.py) that the tracer can lower is traced into a device program
(thermca_b200/trace.py) and stored together with sample temperature pairs and the values the UNMODIFIED
reference closure returns for them.  The known answers of docs/examples/heat_transfer_on_skin.ipynb
(cells at :61,:103,:140,:170,:209) are part of the samples and asserted here.

Usage:  python oracle/make_golden_links.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import refenv  # noqa: E402

refenv.activate()
from thermca import free_conv, forced_conv, radiation, combd_film, evaporation, bearing, fluids, TRANSVERSE  # noqa: E402

from thermca_b200.trace import trace, TableRegistry, UntraceableError  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
air, water = fluids.air, fluids.water

KAT_EMIS = radiation.res_emis(emis0=.95, emis1=.8, view0=1, view1=1)
assert KAT_EMIS == 0.7676767676767677                      # heat_transfer_on_skin.ipynb:61

CASES = [
    ("radiation.therm_radn[kat]", radiation.therm_radn(res_emis=KAT_EMIS), None, 4.8055738452682135),
    ("free_conv.vert_surf[kat]", free_conv.vert_surf(surf_hgt=.25), air, 3.889609859462753),
    ("forced_conv.cyl_cross_flow[kat]", forced_conv.cyl_cross_flow(vel=.1, rad=.035), air, 3.8022233532323804),
    ("combd_film.mix_conv[kat]", combd_film.mix_conv(free_conv.vert_surf(surf_hgt=.25),
                                                     forced_conv.cyl_cross_flow(vel=.1, rad=.035), flow=TRANSVERSE),
     air, 4.574476631734567),
    ("free_conv.vert_surf", free_conv.vert_surf(), air, None),
    ("free_conv.vert_surf[water]", free_conv.vert_surf(surf_hgt=0.2), water, None),
    ("free_conv.horiz_top_surf", free_conv.horiz_top_surf(), air, None),
    ("free_conv.vert_gap", free_conv.vert_gap(), None, None),
    ("free_conv.horiz_gap", free_conv.horiz_gap(), None, None),
    ("free_conv.horiz_cyl_gap", free_conv.horiz_cyl_gap(), None, None),
    ("free_conv.horiz_cyl", free_conv.horiz_cyl(), air, None),
    ("forced_conv.pipe", forced_conv.pipe(), water, None),
    ("forced_conv.rot_cyl_in_air", forced_conv.rot_cyl_in_air(rot_freq=20.0, rad=0.1), air, None),
    ("forced_conv.rot_disc_in_air", forced_conv.rot_disc_in_air(rot_freq=20.0, rad=0.1), air, None),
    ("forced_conv.rot_cone_in_air", forced_conv.rot_cone_in_air(rot_freq=20.0, rad=0.1, cone_angle=0.5), air, None),
    ("forced_conv.rot_cyl_air_gap", forced_conv.rot_cyl_air_gap(rot_freq=20.0, rad=0.1, gap=0.005), air, None),
    ("forced_conv.plate", forced_conv.plate(), air, None),
    ("forced_conv.cyl_cross_flow", forced_conv.cyl_cross_flow(), air, None),
    ("radiation.therm_radn", radiation.therm_radn(), None, None),
    ("radiation.therm_radn_room", radiation.therm_radn_room(), None, None),
    ("combd_film.mix_conv[assisting]", combd_film.mix_conv(free_conv.vert_surf(), forced_conv.plate()), air, None),
    ("combd_film.mix_conv[scalar]", combd_film.mix_conv(free_conv.vert_surf(), 4.0), air, None),
    ("combd_film.conv_radn_room_at_room_temps", combd_film.conv_radn_room_at_room_temps(), air, None),
    ("combd_film.conv_radn_room", combd_film.conv_radn_room(), air, None),
    ("evaporation.free_forc_water_to_air", evaporation.free_forc_water_to_air(), air, None),
    # Python control flow on data (traced path by path and joined with WHERE):
    ("free_conv.vert_cyl_gap", free_conv.vert_cyl_gap(), None, None),              # per-element if/elif loop :317-323
    ("bearing.heat_loss", bearing.heat_loss(), "heat", None),                      # `if v < 2e-4` bearing.py:131
    ("bearing.heat_loss[slow, oil bath]", bearing.heat_loss(rot_freq=0.02, lube_type="oil bath"), "heat", None),
]


def main():
    rng = np.random.default_rng(3)
    t0 = np.concatenate([[36.0], rng.uniform(25.0, 120.0, 47)])
    t1 = np.concatenate([[22.0], rng.uniform(-5.0, 24.0, 47)])
    tables = TableRegistry()
    out = {}
    k = 0
    names = []
    for name, func, matl, kat in CASES:
        heat = isinstance(matl, str) and matl == "heat"          # heat functions take the mean temperature only
        try:
            prog = trace(func, 1, [], tables) if heat else trace(func, 2, [], tables, extra_args=(matl,))
        except UntraceableError as e:
            print("untraceable:", name, e)
            continue
        with np.errstate(all="ignore"):
            if heat:      # network.py:788-790 calls f(mean(T[idx])) with a scalar
                expect = np.array([float(func(a)) for a in t0])
            else:         # array call, as the solver issues it (network.py:781-785: one vector per link group)
                expect = np.asarray(func(t0.copy(), t1.copy(), matl), dtype=np.float64) * np.ones_like(t0)
        if kat is not None:
            assert float(func(36, 22, matl)) == kat, (name, kat)       # the notebook's printed digits (scalar call)
            assert abs(expect[0] - kat) <= 4e-16 * kat                 # numpy's array pow may differ in the last bit
        out[f"{k}_code"] = np.array(prog.code, dtype=np.int32)
        out[f"{k}_consts"] = np.array(prog.consts, dtype=np.float64)
        out[f"{k}_result"] = np.int64(prog.result)
        out[f"{k}_expect"] = expect
        names.append(name)
        print(k, name, len(prog.code), "instructions", "KAT ok" if kat is not None else "")
        k += 1
    out["names"] = np.array(names)
    out["t0"], out["t1"] = t0, t1
    out["n_tables"] = np.int64(len(tables.tables))
    for i, (x, y) in enumerate(tables.tables):
        out[f"table{i}_x"], out[f"table{i}_y"] = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "link_programs.npz"), **out)
    print(k, "programs,", len(tables.tables), "tables")


if __name__ == "__main__":
    main()
