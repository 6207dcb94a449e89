"""Models built through the REAL reference API, for golden-vector generation.

TEST INFRASTRUCTURE ONLY; runs only in the build container (needs /root/reference).
Each builder returns ``(net, sim_kwargs)``.  Sources of the model definitions:
docs/examples/getting_started.ipynb, coffeecup.ipynb (BASELINE config #1),
cutting_disc.ipynb, thermca/tests/test_fe_part.py:102-166 (plate with heat + film).
"""
from __future__ import annotations

from math import tau

import numpy as np

import refenv

refenv.activate()
from thermca import *  # noqa: E402,F401,F403
from thermca import Mesh, Model, Network  # noqa: E402

PLATE_MSH = refenv.REFERENCE_ROOT + "/thermca/tests/plate.msh"
test_steel = Solid(dens=8000.0, spec_heat=500.0, condy=50.0, name="test_steel")


def getting_started():
    with Model() as model:
        body = Node(capy=1.0, init_temp=0.0, name="body")
        bound = BoundNode(temp=1.0)
        CondLink(body, bound, cond=1.0)
    return Network(model), dict(time_span=[0.0, 5.0])


def coffeecup():
    cup_height, cup_outer_rad, cup_wall_thkns = 0.095, 0.0395, 0.0049
    cup_inner_rad = cup_outer_rad - cup_wall_thkns
    air_fill_height = 0.012
    coffee_height = cup_height - cup_wall_thkns - air_fill_height
    with Asm() as asm:
        base = Cyl(outer_rad=cup_inner_rad, lgth=cup_wall_thkns, name="base")
        wet_wall = Cyl(posn=(cup_wall_thkns, 0.0, 0.0), inner_rad=cup_inner_rad,
                       outer_rad=cup_outer_rad, lgth=coffee_height, rad_div=2, name="wet_wall")
        dry_wall = Cyl(posn=(cup_wall_thkns + coffee_height, 0.0, 0.0), inner_rad=cup_inner_rad,
                       outer_rad=cup_outer_rad, lgth=air_fill_height, rad_div=2, name="dry_wall")
        edge = Cyl(inner_rad=cup_inner_rad, outer_rad=cup_outer_rad, lgth=cup_wall_thkns,
                   rad_div=2, name="edge")
        Surf(name="wet", faces=[base.face.end, wet_wall.face.inner])
        Surf(name="dry", faces=[base.face.base, edge.face.base, edge.face.outer, wet_wall.face.outer,
                                dry_wall.face.outer, dry_wall.face.end, dry_wall.face.inner])
    air_temp, init_coffee_temp = 25.0, 100.0
    with Model() as model:
        cup = LPPart(asm=asm, matl=solids.porcelain, init_temp=air_temp + 0.1, name="cup")
        coffee = MatlNode(matl=fluids.water, init_temp=init_coffee_temp,
                          vol=tau / 2.0 * cup_inner_rad ** 2 * coffee_height,
                          posn=(coffee_height / 2.0 + cup_wall_thkns, 0.0, 0.0), name="coffee")
        evap_sink = BoundNode(temp=air_temp, posn=(0.1, 0.015, 0.0), name="evaporation_sink")
        conv_sink = BoundNode(posn=(0.02, 0.055, 0.0), temp=air_temp, name="convection_sink")
        radn_sink = BoundNode(temp=air_temp, posn=(0.07, 0.055, 0.0), name="radiation_sink")
        free_conv_cup = free_conv.vert_surf(surf_hgt=cup_height)
        FilmLink(cup.surf.wet, coffee, free_conv_cup, matl=fluids.water)
        FilmLink(cup.surf.dry, conv_sink, free_conv_cup, matl=fluids.air)
        emis_room = 0.8
        FilmLink(cup.surf.dry, radn_sink, radiation.therm_radn(res_emis=0.8 * emis_room))
        area = tau / 2.0 * cup_inner_rad ** 2.0
        CondLink(coffee, conv_sink,
                 free_conv.horiz_top_surf(surf_width=2 * cup_inner_rad, factor=area), matl=fluids.air)
        CondLink(coffee, radn_sink, radiation.therm_radn(res_emis=0.8 * emis_room, factor=area))
        CondLink(coffee, evap_sink,
                 evaporation.free_forc_water_to_air(vel=0.0, rel_hum=0.6, factor=area))
    net = Network(model)
    net._golden_probe = {"coffee": int(net.elem_idx[coffee][0]), "wet": int(net.elem_idx[cup.surf.wet][0])}
    return net, dict(time_span=[0.0, 1200.0], rel_tol=1e-5)


def cutting_disc():
    with Asm() as disc_asm:
        cyl = Cyl(inner_rad=0.011, outer_rad=0.075, lgth=0.0015, rad_div=16)
        Surf(name="circ_faces", faces=[cyl.face.base, cyl.face.end])
        Surf(name="outer", faces=[cyl.face.outer])
        Surf(name="inner", faces=[cyl.face.inner])
    with Model() as model:
        abrasive = Solid.mix(matls=(solids.alumina, solids.glass), shares=(0.9, 0.1))
        disc = LPPart(asm=disc_asm, matl=abrasive, init_temp=20.0, name="flex_disc")
        env_temp = Input([[0, 19], [20, 23], [40, 27]], name="environment_temperatures")
        environment = BoundNode(temp=env_temp.get_value, name="environment")
        rpm = tau / 60
        rot_freq = Input([[0, 3000 * rpm], [20, 6000 * rpm]])
        FilmLink(disc.surf.circ_faces, environment,
                 film=forced_conv.rot_disc_in_air(rot_freq=rot_freq.value, rad=cyl.outer_rad))
        cutting_heat = Input([[0, 40], [20, 80]])
        HeatSource(disc.surf.outer, heat=lambda temp: cutting_heat.value)
    net = Network(model)
    net._golden_probe = {"outer": int(net.elem_idx[disc.surf.outer][0]),
                         "inner": int(net.elem_idx[disc.surf.inner][0])}
    return net, dict(time_span=[0.0, 50.0])


def _plate(mor_dof, variable):
    mesh = Mesh.read(PLATE_MSH, file_format="gmsh")
    with Model() as model:
        plate = FEPart(mesh, test_steel, init_temp=20.0, mor_dof=mor_dof, lump_cond_meth=None,
                       name="plate")
        env = BoundNode(temp=20.0, name="environment")
        HeatSource(plate.surf.bottom, 1000.0)
        if not variable:
            FilmLink.multi([plate.surf.top, plate.surf.left, plate.surf.right,
                            plate.surf.front, plate.surf.back], env, 10.0)
        else:
            # nonlinear films, an unknown network node, a stationary node and an Input
            FilmLink(plate.surf.top, env, free_conv.vert_surf(surf_hgt=0.5), matl=fluids.air)
            FilmLink(plate.surf.left, env, radiation.therm_radn(res_emis=0.8))
            lump = Node(capy=2.0e3, init_temp=25.0, name="lump")
            FilmLink(plate.surf.right, lump, 40.0)
            FilmLink(plate.surf.right, env, 5.0)        # two films on one surface
            stat = StatNode(name="stat")
            CondLink(lump, stat, 3.0)
            CondLink(stat, env, combd_film.conv_radn_room(factor=0.5), matl=fluids.air)
            power = Input([[0, 0.0], [30, 150.0], [90, 50.0]])
            HeatSource(lump, heat=lambda temp: power.value * (1.0 + 0.0 * temp))
            FluxSource(plate.surf.front, flux=lambda temp: 200.0 - 2.0 * temp)
    return Network(model)


def plate_ffe():
    return _plate(None, False), dict(time_span=[0.0, 30.0], method="RK45")


def plate_mor():
    return _plate(10, False), dict(time_span=[0.0, 600.0], method="RK45")


def plate_ffe_var():
    return _plate(None, True), dict(time_span=[0.0, 20.0], method="RK45")


def plate_mor_var():
    return _plate(12, True), dict(time_span=[0.0, 300.0], method="RK23")


def kuhn_cuboid():
    """Small instance of the synthetic BASELINE config #2 generator, through the real API."""
    from thermca_b200.synth import kuhn_cuboid_mesh
    pts, tets, tri_bottom, tri_upper = kuhn_cuboid_mesh(6, 12, 3, 0.5, 1.0, 0.05)
    mesh = Mesh(pts, [tets, tri_bottom, tri_upper], ["tetra", "triangle", "triangle"],
                ["cuboid", "bottom", "upper_faces"])
    with Model() as model:
        cuboid = FEPart(mesh, test_steel, init_temp=20.0, lump_cond_meth=None, name="cuboid")
        env = BoundNode(temp=20.0, name="environment")
        HeatSource(cuboid.surf.bottom, 1000.0)
        FilmLink(cuboid.surf.upper_faces, env, 10.0)
    return Network(model), dict(time_span=[0.0, 5.0], method="RK45")


def tdep_rod():
    """Temperature dependent conductivity and density on a synthetic rod
    (pattern of thermca/tests/test_fe_part.py:572-661; the .med meshes are unreadable here)."""
    from thermca_b200.synth import kuhn_cuboid_mesh
    pts, tets, tri_left, tri_rest = kuhn_cuboid_mesh(10, 1, 1, 1.0, 0.05, 0.05)
    # surfaces: x=0 face ('left'), x=L face ('right')
    x = pts[:, 0]
    allt = np.vstack([tri_left, tri_rest])
    left = allt[np.all(np.isclose(x[allt], 0.0), axis=1)]
    right = allt[np.all(np.isclose(x[allt], 1.0), axis=1)]
    mesh = Mesh(pts, [tets, left, right], ["tetra", "triangle", "triangle"], ["rod", "left", "right"])
    matl = Solid(condy=np.array([[0.0, 5.0], [0.1, 0.5]]), dens=np.array([[0.0, 5.0], [1.0, 0.5]]),
                 spec_heat=2.0)
    with Model() as model:
        rod = FEPart(mesh=mesh, matl=matl, init_temp=1.0, temp_dependent=True, lump_cond_meth=None,
                     name="rod")
        env = BoundNode(temp=0.0)
        FluxSource(rod.surf.left, flux=1.0)
        FilmLink(rod.surf.right, env, 1.0e3)
    return Network(model), dict(time_span=[0.01, 0.3], method="RK45")


def assembly():
    """Small instance of BASELINE config #4: FE bodies (one temperature dependent, one MOR), two
    linked LP blocks, temperature dependent fluid nodes with flow links, a stationary node,
    nonlinear films, an Input driven heat loss (pattern of docs/examples/bearing.ipynb)."""
    mesh = Mesh.read(PLATE_MSH, file_format="gmsh")
    with Model() as model:
        frame = FEPart(mesh, solids.steel, init_temp=22.0, temp_dependent=True, lump_cond_meth=None,
                       name="frame")
        table = FEPart(mesh, test_steel, init_temp=21.0, mor_dof=8, lump_cond_meth=None, name="table")
        blk1 = lp_parts.block(matl=solids.steel, init_temp=23.0, width=0.2, hgt=0.1, depth=0.1,
                              width_div=4, hgt_div=2, depth_div=2, name="blk1")
        blk2 = lp_parts.block(matl=solids.al, init_temp=20.0, width=0.1, hgt=0.1, depth=0.1,
                              width_div=2, hgt_div=2, depth_div=2, name="blk2")
        env = BoundNode(temp=20.0, name="env")
        inlet = BoundNode(temp=30.0, name="inlet")
        oil1 = MatlNode(vol=2e-4, matl=fluids.water, init_temp=25.0, name="oil1")
        oil2 = MatlNode(vol=3e-4, matl=fluids.water, init_temp=24.0, name="oil2")
        oil1.temp_dependent = True
        oil2.temp_dependent = True
        FlowLink(oil1, inlet, vol_flow=1.0e-5)
        FlowLink(oil2, oil1, vol_flow=lambda t_dest, t_src, fluid: 1.0e-5 * (1.0 + 0.01 * (t_src - 20.0)),
                 matl=fluids.water)
        FilmLink(frame.surf.top, oil1, forced_conv.plate(vel=0.5, length=0.5), matl=fluids.water)
        FilmLink(table.surf.top, oil2, 200.0)
        FilmLink(blk1.surf.top, frame.surf.bottom, 500.0)
        FilmLink(blk1.surf.right, blk2.surf.left, film=lambda t0, t1, matl: 300.0 + 0.5 * (t0 + t1))
        FilmLink(blk2.surf.right, env, combd_film.conv_radn_room(), matl=fluids.air)
        FilmLink(frame.surf.left, env, free_conv.vert_surf(surf_hgt=0.05), matl=fluids.air)
        FilmLink(table.surf.left, env, radiation.therm_radn(res_emis=0.7))
        hub = StatNode(name="hub")
        CondLink(hub, oil2, 4.0)
        CondLink(hub, env, 1.5)
        speed = Input([[0, 100.0], [20, 250.0], [40, 150.0]], name="speed")
        HeatSource(table.surf.bottom, heat=lambda temp: 0.5 * speed.value * (1.0 - 0.002 * (temp - 20.0)))
        HeatSource(hub, 12.0)
    return Network(model), dict(time_span=[0.0, 45.0], method="RK45", rel_tol=1e-4)


MODELS = {
    "getting_started": getting_started,
    "coffeecup": coffeecup,
    "cutting_disc": cutting_disc,
    "plate_ffe": plate_ffe,
    "plate_mor": plate_mor,
    "plate_ffe_var": plate_ffe_var,
    "plate_mor_var": plate_mor_var,
    "kuhn_cuboid": kuhn_cuboid,
    "tdep_rod": tdep_rod,
    "assembly": assembly,
}


def lp_tdep():
    """Temperature dependent LP body (pattern of thermca/tests/test_lp_part.py:601-665, shorter rod)."""
    with Asm() as rod_asm:
        cyl_rod = Cyl(lgth=1.0, lgth_div=24, name="cyl")
        Surf(name="begin", faces=[cyl_rod.face.base])
        Surf(name="end", faces=[cyl_rod.face.end])
    matl = Solid(condy=np.array([[0.0, 5.0], [0.1, 0.5]]), dens=np.array([[0.0, 5.0], [1.0, 0.8]]), spec_heat=2.0)
    with Model() as model:
        rod = LPPart(asm=rod_asm, matl=matl, init_temp=1.0, temp_dependent=True, name="rod")
        stat = StatNode()
        FilmLink(rod.surf.begin, stat, film=1e32)
        HeatSource(stat, heat=1.0 * rod.surf.begin.area())
        bound = BoundNode(temp=0.0)
        FilmLink(rod.surf.end, bound, film=free_conv.vert_surf(surf_hgt=0.3, factor=50.0), matl=fluids.air)
    return Network(model), dict(time_span=[0.01, 3.0], method="RK45")


MODELS["lp_tdep"] = lp_tdep


def _plate_stat(mor_dof, variable):
    """Plate fed and cooled through stationary nodes (the reference's idiom for a prescribed surface heat flow:
    StatNode + HeatSource + a stiff film, thermca/tests/test_lp_part.py:655-660 — here on an FE / MOR body).
    Exercises the W-methods' Jacobian model for films that lead to stationary nodes (csrc/network.cuh C_JFAC)."""
    mesh = Mesh.read(PLATE_MSH, file_format="gmsh")
    with Model() as model:
        plate = FEPart(mesh, test_steel, init_temp=20.0, mor_dof=mor_dof, lump_cond_meth=None, name="plate")
        env = BoundNode(temp=20.0, name="environment")
        feed = StatNode(name="feed")
        FilmLink(plate.surf.bottom, feed, 1.0e5)                # the node only carries the source: flux condition
        HeatSource(feed, heat=1000.0)
        mid = StatNode(name="mid")                              # film in series with a conductance
        if variable:
            FilmLink(plate.surf.top, mid, free_conv.vert_surf(surf_hgt=0.5), matl=fluids.air)
        else:
            FilmLink(plate.surf.top, mid, 2000.0)
        CondLink(mid, env, 2.0)
        FilmLink.multi([plate.surf.left, plate.surf.right], env, 10.0)
    return Network(model)


def plate_ffe_stat():
    return _plate_stat(None, True), dict(time_span=[0.0, 30.0], method="RK45")


def plate_ffe_statc():
    return _plate_stat(None, False), dict(time_span=[0.0, 30.0], method="RK45")


def plate_mor_stat():
    return _plate_stat(10, True), dict(time_span=[0.0, 300.0], method="RK45")


MODELS["plate_ffe_stat"] = plate_ffe_stat
MODELS["plate_ffe_statc"] = plate_ffe_statc
MODELS["plate_mor_stat"] = plate_mor_stat


def steel_sheet():
    """The LPPart steel sheet of the ``Result`` doctest (thermca/resultdata.py:81-131): its printed
    temperature table is a known answer of the reference for an LP transient."""
    steel = Solid(dens=8000.0, spec_heat=500.0, condy=50.0)
    with Asm() as cuboid:
        block = Cube(width=0.5, hgt=1.0, depth=0.05, hgt_div=10)
        face = block.face
        Surf(name="upper", faces=[face.left, face.right, face.front, face.back, face.top])
        Surf(name="btm", faces=[face.btm])
    with Model() as model:
        sheet = LPPart(asm=cuboid, matl=steel, name="sheet")
        env = BoundNode(temp=20.0, posn=(-0.3, 0.5, 0.025), name="env")
        FilmLink(sheet.surf.upper, env, film=10.0)
        HeatSource(sheet.surf.btm, heat=1000.0, name="src")
    return Network(model), dict(time_span=[0.0, 36000.0], method="RK45")


MODELS["steel_sheet"] = steel_sheet


def _cuboid_mesh_named(nx, ny, nz, lx, ly, lz):
    """Kuhn cuboid through the reference Mesh API with the surfaces top / bottom / left (x=0) / rest."""
    from thermca_b200.synth import kuhn_cuboid_mesh
    pts, tets, tri_bottom, tri_upper = kuhn_cuboid_mesh(nx, ny, nz, lx, ly, lz, jitter=0.15)
    z, x = pts[:, 2], pts[:, 0]
    top = np.all(np.isclose(z[tri_upper], lz), axis=1)
    left = np.all(np.isclose(x[tri_upper], 0.0), axis=1) & ~top
    return Mesh(pts, [tets, tri_upper[top], tri_bottom, tri_upper[left], tri_upper[~top & ~left]],
                ["tetra"] + ["triangle"] * 4, ["body", "top", "bottom", "left", "rest"])


def assembly_large(frame_cells=(36, 72, 8), table_cells=(24, 48, 6)):
    """BASELINE config #4 at its stated size: FE bodies of 10^4..10^5 DOF (one temperature dependent full-order
    body, one MOR body), two linked LP blocks, temperature dependent fluid nodes with flow links, a stationary node,
    nonlinear films, an Input driven heat loss — the pattern of ``assembly`` (docs/examples/bearing.ipynb with FE
    bodies substituted) on meshes from the synthetic generator."""
    frame_mesh = _cuboid_mesh_named(*frame_cells, 0.5, 1.0, 0.05)
    table_mesh = _cuboid_mesh_named(*table_cells, 0.4, 0.8, 0.04)
    with Model() as model:
        frame = FEPart(frame_mesh, solids.steel, init_temp=22.0, temp_dependent=True, lump_cond_meth=None,
                       name="frame")
        table = FEPart(table_mesh, test_steel, init_temp=21.0, mor_dof=8, lump_cond_meth=None, name="table")
        blk1 = lp_parts.block(matl=solids.steel, init_temp=23.0, width=0.2, hgt=0.1, depth=0.1,
                              width_div=4, hgt_div=2, depth_div=2, name="blk1")
        blk2 = lp_parts.block(matl=solids.al, init_temp=20.0, width=0.1, hgt=0.1, depth=0.1,
                              width_div=2, hgt_div=2, depth_div=2, name="blk2")
        env = BoundNode(temp=20.0, name="env")
        inlet = BoundNode(temp=30.0, name="inlet")
        oil1 = MatlNode(vol=2e-4, matl=fluids.water, init_temp=25.0, name="oil1")
        oil2 = MatlNode(vol=3e-4, matl=fluids.water, init_temp=24.0, name="oil2")
        oil1.temp_dependent = True
        oil2.temp_dependent = True
        FlowLink(oil1, inlet, vol_flow=1.0e-5)
        FlowLink(oil2, oil1, vol_flow=lambda t_dest, t_src, fluid: 1.0e-5 * (1.0 + 0.01 * (t_src - 20.0)),
                 matl=fluids.water)
        FilmLink(frame.surf.top, oil1, forced_conv.plate(vel=0.5, length=0.5), matl=fluids.water)
        FilmLink(table.surf.top, oil2, 200.0)
        FilmLink(blk1.surf.top, frame.surf.bottom, 500.0)
        FilmLink(blk1.surf.right, blk2.surf.left, film=lambda t0, t1, matl: 300.0 + 0.5 * (t0 + t1))
        FilmLink(blk2.surf.right, env, combd_film.conv_radn_room(), matl=fluids.air)
        FilmLink(frame.surf.left, env, free_conv.vert_surf(surf_hgt=0.05), matl=fluids.air)
        FilmLink(table.surf.left, env, radiation.therm_radn(res_emis=0.7))
        hub = StatNode(name="hub")
        CondLink(hub, oil2, 4.0)
        CondLink(hub, env, 1.5)
        speed = Input([[0, 100.0], [20, 250.0], [40, 150.0]], name="speed")
        HeatSource(table.surf.bottom, heat=lambda temp: 0.5 * speed.value * (1.0 - 0.002 * (temp - 20.0)))
        HeatSource(hub, 12.0)
    return Network(model), dict(time_span=[0.0, 20.0], method="RK45", rel_tol=1e-4)
