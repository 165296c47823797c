"""Tuning harness (development aid): build variants of ONE unit with different tiling / role / roll settings into small
libraries (tune_libs/, git-ignored), then time them on the GPU.

    python scripts/tune.py build  <workload>      # here (no GPU): compiles all variants in parallel
    python scripts/tune.py run    <workload>      # on the GPU box: times every tune_libs/<workload>_*.so and checks parity
"""
import concurrent.futures as cf
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "robotdynamics.jl_b200", "csrc")
OUT = os.path.join(ROOT, "tune_libs")
sys.path.insert(0, ROOT)
sys.path.insert(0, CSRC)

UNIT = {   # workload -> (unit name, -D flags, bench.py workload)
    "quadrotor": ("quad_quat_world_f32", dict(RDB_KIND=1, RDB_ROT=1, RDB_FRAME=0, RDB_DTYPE=0)),
    "cartpole": ("cartpole_f64", dict(RDB_KIND=0, RDB_DTYPE=1)),
    "quaderr": ("quad_quat_world_f32", dict(RDB_KIND=1, RDB_ROT=1, RDB_FRAME=0, RDB_DTYPE=0)),
    "bodyquat": ("body_quat_world_f32", dict(RDB_KIND=2, RDB_ROT=1, RDB_FRAME=0, RDB_DTYPE=0)),
    "quadmrp": ("quad_mrp_world_f32", dict(RDB_KIND=1, RDB_ROT=2, RDB_FRAME=0, RDB_DTYPE=0)),
    "quadbody": ("quad_quat_body_f32", dict(RDB_KIND=1, RDB_ROT=1, RDB_FRAME=1, RDB_DTYPE=0)),
    "quadrotor64": ("quad_quat_world_f64", dict(RDB_KIND=1, RDB_ROT=1, RDB_FRAME=0, RDB_DTYPE=1)),
    "satellite": ("body_mrp_world_f64", dict(RDB_KIND=2, RDB_ROT=2, RDB_FRAME=0, RDB_DTYPE=1)),
    "satellite32": ("body_mrp_world_f32", dict(RDB_KIND=2, RDB_ROT=2, RDB_FRAME=0, RDB_DTYPE=0)),
    "quadbody64": ("quad_quat_body_f64", dict(RDB_KIND=1, RDB_ROT=1, RDB_FRAME=1, RDB_DTYPE=1)),
    "quadmrp64": ("quad_mrp_world_f64", dict(RDB_KIND=1, RDB_ROT=2, RDB_FRAME=0, RDB_DTYPE=1)),
}

VARIANTS = {
    "quadrotor": {
        "base": {},
        "split13": dict(RDB_TUNE_C0="0x1FFFu", RDB_TUNE_C1="0x1E000u"),
        "t64_minb2": dict(RDB_TUNE_TILE=64, RDB_TUNE_MINB=2),
        "t32_minb4": dict(RDB_TUNE_TILE=32, RDB_TUNE_MINB=4),
        "t96_minb1": dict(RDB_TUNE_TILE=96, RDB_TUNE_MINB=1),
    },
    "quaderr": {
        "base": {},
        "rowbulk": dict(RDB_TUNE_JMAP=0),
        "dense": dict(RDB_TUNE_ROWSTORE=0),
        "3r": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0x3Fu", RDB_TUNE_C1="0xFC0u", RDB_TUNE_C2="0xF000u"),
        "2r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0x7FFu", RDB_TUNE_C1="0xF800u"),
        "2rb": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1FFu", RDB_TUNE_C1="0xFE00u"),
        "unroll": dict(RDB_TUNE_ROLL=0),
    },
    "quadbody": {
        "base": {},
        "roll2_2r": dict(RDB_TUNE_ROLL=2),
        "roll2_3r": dict(RDB_TUNE_ROLL=2, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x1F80u", RDB_TUNE_C2="0x1E000u"),
        "roll1_3rc": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x3FFu", RDB_TUNE_C1="0x1C00u", RDB_TUNE_C2="0x1E000u"),
        "roll1_2rc": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x3FFu", RDB_TUNE_C1="0x1FC00u"),
    },
    "bodyquat": {
        "base": {},
        "2r_a": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0xFFFu", RDB_TUNE_C1="0x7F000u"),
        "2r_b": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x3FFu", RDB_TUNE_C1="0x7FC00u"),
        "3r_t128": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x3FFu", RDB_TUNE_C1="0x7C00u", RDB_TUNE_C2="0x78000u"),
        "3r_b": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0xFFFu", RDB_TUNE_C1="0xF000u", RDB_TUNE_C2="0x70000u"),
    },
    "quadmrp": {
        "base": {},
        "rowbulk": dict(RDB_TUNE_JMAP=0),
        "3r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0x3Fu", RDB_TUNE_C1="0xFC0u", RDB_TUNE_C2="0xF000u"),
        "2rb": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1FFu", RDB_TUNE_C1="0xFE00u"),
        "roll2": dict(RDB_TUNE_ROLL=2),
    },
    "quadrotor64": {
        "base": {},
        "4rb": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x3Fu", RDB_TUNE_C1="0x7C0u", RDB_TUNE_C2="0x3800u", RDB_TUNE_C3="0x1C000u"),
        "4rc": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0xF80u", RDB_TUNE_C2="0x7000u", RDB_TUNE_C3="0x18000u"),
        "5rb": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1Fu", RDB_TUNE_C1="0x3E0u", RDB_TUNE_C2="0x1C00u", RDB_TUNE_C3="0x6000u", RDB_TUNE_C4="0x18000u"),
        "3rb": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0xFFu", RDB_TUNE_C1="0x1F00u", RDB_TUNE_C2="0x1E000u"),
    },
    "cartpole": {
        "base": {},
        "t32_minb16": dict(RDB_TUNE_TILE=32, RDB_TUNE_MINB=16),
        "t32_minb14": dict(RDB_TUNE_TILE=32, RDB_TUNE_MINB=14),
        "t32_minb18": dict(RDB_TUNE_TILE=32, RDB_TUNE_MINB=18),
        "t64_minb8": dict(RDB_TUNE_TILE=64, RDB_TUNE_MINB=8),
        "t64_minb7": dict(RDB_TUNE_TILE=64, RDB_TUNE_MINB=7),
        "roll1_minb16": dict(RDB_TUNE_TILE=32, RDB_TUNE_MINB=16, RDB_TUNE_ROLL=1),
        "t64_minb9": dict(RDB_TUNE_TILE=64, RDB_TUNE_MINB=9),
        "t96_minb6": dict(RDB_TUNE_TILE=96, RDB_TUNE_MINB=6),
        "t128_minb4": dict(RDB_TUNE_TILE=128, RDB_TUNE_MINB=4),
    },
    "satellite": {
        "base": {},
        "rowbulk": dict(RDB_TUNE_JMAP=0),
        "t64_3c": dict(RDB_TUNE_TILE=64, RDB_TUNE_MINB=1),
        "c2": dict(RDB_TUNE_C0="0x3u", RDB_TUNE_C1="0xCu", RDB_TUNE_C2="0x30u", RDB_TUNE_C3="0xC0u", RDB_TUNE_C4="0x300u", RDB_TUNE_C5="0xC00u", RDB_TUNE_C6="0x3000u", RDB_TUNE_C7="0xC000u", RDB_TUNE_C8="0x30000u", RDB_TUNE_TILE=32, RDB_TUNE_MINB=2),
        "c6": dict(RDB_TUNE_C0="0x3Fu", RDB_TUNE_C1="0xFC0u", RDB_TUNE_C2="0x3F000u", RDB_TUNE_TILE=64, RDB_TUNE_MINB=2),
        "c9": dict(RDB_TUNE_C0="0x1FFu", RDB_TUNE_C1="0x3FE00u", RDB_TUNE_TILE=64, RDB_TUNE_MINB=2),
        "c18": dict(RDB_TUNE_C0="0x3FFFFu", RDB_TUNE_TILE=64, RDB_TUNE_MINB=2),
    },
}
# round 2: packed fp32x2 partial arithmetic on top of the elemental rotations (dense chains), and role splits for the fp64 kernels
for _w in ("quadrotor", "quadbody", "quadmrp", "bodyquat"):
    VARIANTS[_w] = {"base": {}, "pack": dict(RDB_PACK_F32=1)}
VARIANTS["quadbody"].update({"3r": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x1F80u", RDB_TUNE_C2="0x1E000u"),
                             "pack_3r": dict(RDB_PACK_F32=1, RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x1F80u", RDB_TUNE_C2="0x1E000u")})
VARIANTS["quadbody64"] = {
    "base": {},
    "6r_t32": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7u", RDB_TUNE_C1="0x78u", RDB_TUNE_C2="0x380u", RDB_TUNE_C3="0x1C00u", RDB_TUNE_C4="0x6000u", RDB_TUNE_C5="0x18000u"),
    "6r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7u", RDB_TUNE_C1="0x78u", RDB_TUNE_C2="0x380u", RDB_TUNE_C3="0x1C00u", RDB_TUNE_C4="0x6000u", RDB_TUNE_C5="0x18000u"),
    "3r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x1F80u", RDB_TUNE_C2="0x1E000u"),
    "4r_t32": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=2),
    "unroll_4r": dict(RDB_TUNE_ROLL=0),
}
VARIANTS["quadmrp64"] = {
    "base": {},
    "6r_t32": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7u", RDB_TUNE_C1="0x38u", RDB_TUNE_C2="0x1C0u", RDB_TUNE_C3="0xE00u", RDB_TUNE_C4="0x3000u", RDB_TUNE_C5="0xC000u"),
    "3r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x3Fu", RDB_TUNE_C1="0xFC0u", RDB_TUNE_C2="0xF000u"),
    "4r_t32": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=2),
}
VARIANTS["cartpole"].update({"null": dict(RDB_TUNE_NULLMODEL=1), "null_t128": dict(RDB_TUNE_NULLMODEL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=4),
                             "null_t32": dict(RDB_TUNE_NULLMODEL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=16),
                             "near8": dict(RDB_TUNE_NEAR8=1), "nopf": dict(RDB_TUNE_NO_PREFETCH=1), "nostream": dict(RDB_TUNE_NO_STREAMOUT=1),
                             "nopf_nostream": dict(RDB_TUNE_NO_PREFETCH=1, RDB_TUNE_NO_STREAMOUT=1), "t128_minb4": dict(RDB_TUNE_TILE=128, RDB_TUNE_MINB=4),
                             "t128_minb3": dict(RDB_TUNE_TILE=128, RDB_TUNE_MINB=3)})
VARIANTS["cartpole"].update({"dense": dict(RDB_TUNE_ROWSTORE=0), "dense_nostream": dict(RDB_TUNE_ROWSTORE=0, RDB_TUNE_NO_STREAMOUT=1),
                             "dense_t128": dict(RDB_TUNE_ROWSTORE=0, RDB_TUNE_TILE=128, RDB_TUNE_MINB=4), "dense_t256": dict(RDB_TUNE_ROWSTORE=0, RDB_TUNE_TILE=256, RDB_TUNE_MINB=2),
                             "t256": dict(RDB_TUNE_TILE=256, RDB_TUNE_MINB=2), "null_dense": dict(RDB_TUNE_NULLMODEL=1, RDB_TUNE_ROWSTORE=0),
                             "null_dense_t128": dict(RDB_TUNE_NULLMODEL=1, RDB_TUNE_ROWSTORE=0, RDB_TUNE_TILE=128, RDB_TUNE_MINB=4),
                             "null_dense_t256": dict(RDB_TUNE_NULLMODEL=1, RDB_TUNE_ROWSTORE=0, RDB_TUNE_TILE=256, RDB_TUNE_MINB=2),
                             "null_dense_nostream": dict(RDB_TUNE_NULLMODEL=1, RDB_TUNE_ROWSTORE=0, RDB_TUNE_NO_STREAMOUT=1)})
VARIANTS["cartpole"].update({"estrin": dict(RDB_TUNE_ESTRIN=1), "estrin_dense_t128": dict(RDB_TUNE_ESTRIN=1, RDB_TUNE_ROWSTORE=0, RDB_TUNE_TILE=128, RDB_TUNE_MINB=4),
                             "near8_dense_t128": dict(RDB_TUNE_NEAR8=1, RDB_TUNE_ROWSTORE=0, RDB_TUNE_TILE=128, RDB_TUNE_MINB=4),
                             "near8_dense": dict(RDB_TUNE_NEAR8=1, RDB_TUNE_ROWSTORE=0), "estrin_dense": dict(RDB_TUNE_ESTRIN=1, RDB_TUNE_ROWSTORE=0),
                             "near8_t128": dict(RDB_TUNE_NEAR8=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=4)})
for _w in ("quadrotor", "satellite"):
    VARIANTS[_w].update({"nopf": dict(RDB_TUNE_NO_PREFETCH=1), "nostream": dict(RDB_TUNE_NO_STREAMOUT=1), "nopf_nostream": dict(RDB_TUNE_NO_PREFETCH=1, RDB_TUNE_NO_STREAMOUT=1)})
VARIANTS["cartpole_value"] = {"base": {}, "t128_minb8": dict(RDB_TUNE_V_TILE=128, RDB_TUNE_V_MINB=8), "t64_minb16": dict(RDB_TUNE_V_TILE=64, RDB_TUNE_V_MINB=16),
                              "t64_minb12": dict(RDB_TUNE_V_TILE=64, RDB_TUNE_V_MINB=12), "t32_minb24": dict(RDB_TUNE_V_TILE=32, RDB_TUNE_V_MINB=24), "t256_minb4": dict(RDB_TUNE_V_TILE=256, RDB_TUNE_V_MINB=4)}
UNIT["cartpole_value"] = UNIT["cartpole"]
VARIANTS["quadrotor_value"] = {"base": {}, "t128_minb8": dict(RDB_TUNE_V_TILE=128, RDB_TUNE_V_MINB=8), "t64_minb12": dict(RDB_TUNE_V_TILE=64, RDB_TUNE_V_MINB=12)}
UNIT["quadrotor_value"] = UNIT["quadrotor"]
VARIANTS["satellite32"] = {"base": {}, "pad8": dict(RDB_ROWSTORE_MINWAY=8), "c9": dict(RDB_TUNE_C0="0x1FFu", RDB_TUNE_C1="0x3FE00u", RDB_TUNE_TILE=64, RDB_TUNE_MINB=3),
                           "c18": dict(RDB_TUNE_C0="0x3FFFFu", RDB_TUNE_TILE=64, RDB_TUNE_MINB=2)}


# round 2, after the split body-frame force (models.cuh: |q|^4 Fb + q \\ Gw): the balance between the roles moved
UNIT["bodybody"] = ("body_quat_body_f32", dict(RDB_KIND=2, RDB_ROT=1, RDB_FRAME=1, RDB_DTYPE=0))
UNIT["bodybody64"] = ("body_quat_body_f64", dict(RDB_KIND=2, RDB_ROT=1, RDB_FRAME=1, RDB_DTYPE=1))
UNIT["quadmrpbody"] = ("quad_mrp_body_f32", dict(RDB_KIND=1, RDB_ROT=2, RDB_FRAME=1, RDB_DTYPE=0))
VARIANTS["quadbody"] = {
    "base": {},
    "s12": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0xFFFu", RDB_TUNE_C1="0x1F000u"),
    "s9": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1FFu", RDB_TUNE_C1="0x1FE00u"),
    "s11": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7FFu", RDB_TUNE_C1="0x1F800u"),
    "3r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x1F80u", RDB_TUNE_C2="0x1E000u"),
    "3r_t64_m1": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x1F80u", RDB_TUNE_C2="0x1E000u"),
    "unroll": dict(RDB_TUNE_ROLL=0),
}
VARIANTS["bodybody"] = {   # NZ = 19: r 0-2, q 3-6, v 7-9, w 10-12, u 13-18
    "base": {},
    "s12": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0xFFFu", RDB_TUNE_C1="0x7F000u"),
    "s9": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1FFu", RDB_TUNE_C1="0x7FE00u"),
    "3r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x1F80u", RDB_TUNE_C2="0x7E000u"),
}
VARIANTS["quadmrpbody"] = {   # NZ = 16: r 0-2, p 3-5, v 6-8, w 9-11, u 12-15
    "base": {},
    "s11": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7FFu", RDB_TUNE_C1="0xF800u"),
    "s8": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1, RDB_TUNE_C0="0xFFu", RDB_TUNE_C1="0xFF00u"),
    "3r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=2, RDB_TUNE_C0="0x3Fu", RDB_TUNE_C1="0xFC0u", RDB_TUNE_C2="0xF000u"),
}
VARIANTS["quadbody64"] = {
    "base": {},
    "5r_t32": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1Fu", RDB_TUNE_C1="0x1E0u", RDB_TUNE_C2="0xE00u", RDB_TUNE_C3="0x7000u", RDB_TUNE_C4="0x18000u"),
    "4r_b": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x780u", RDB_TUNE_C2="0x3800u", RDB_TUNE_C3="0x1C000u"),
    "4r_c": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1Fu", RDB_TUNE_C1="0x3E0u", RDB_TUNE_C2="0x1C00u", RDB_TUNE_C3="0x1E000u"),
    "3r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x1F80u", RDB_TUNE_C2="0x1E000u"),
    "4r_t32_m2": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=2),
}
VARIANTS["bodybody64"] = {
    "base": {},
    "4r_b": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x780u", RDB_TUNE_C2="0x7800u", RDB_TUNE_C3="0x78000u"),
    "5r_t32": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1Fu", RDB_TUNE_C1="0x1E0u", RDB_TUNE_C2="0x1E00u", RDB_TUNE_C3="0xE000u", RDB_TUNE_C4="0x70000u"),
    "3r_t64": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x7Fu", RDB_TUNE_C1="0x1F80u", RDB_TUNE_C2="0x7E000u"),
}


# small-batch latency (N <= a few thousand knots: one wave, the time is one thread's dependent chain): narrow tiles, more / narrower roles
UNIT["quadlat"] = UNIT["quadrotor"]
VARIANTS["quadlat"] = {
    "base": {},
    "t32_2r": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0xFFFu", RDB_TUNE_C1="0x1F000u"),
    "t32_4r": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x3Fu", RDB_TUNE_C1="0x7C0u", RDB_TUNE_C2="0x3800u", RDB_TUNE_C3="0x1C000u"),
    "t32_6r": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1Fu", RDB_TUNE_C1="0xE0u", RDB_TUNE_C2="0x700u", RDB_TUNE_C3="0x1800u", RDB_TUNE_C4="0x6000u", RDB_TUNE_C5="0x18000u"),
    "t32_8r": dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0xFu", RDB_TUNE_C1="0x30u", RDB_TUNE_C2="0xC0u", RDB_TUNE_C3="0x300u", RDB_TUNE_C4="0xC00u", RDB_TUNE_C5="0x3000u", RDB_TUNE_C6="0xC000u", RDB_TUNE_C7="0x10000u"),
    "t32_6r_unroll": dict(RDB_TUNE_ROLL=0, RDB_TUNE_TILE=32, RDB_TUNE_MINB=1, RDB_TUNE_C0="0x1Fu", RDB_TUNE_C1="0xE0u", RDB_TUNE_C2="0x700u", RDB_TUNE_C3="0x1800u", RDB_TUNE_C4="0x6000u", RDB_TUNE_C5="0x18000u"),
}


# instruction-cache pressure (ncu: no_instruction 0.38 / 0.73 / 1.99 warps per issue for code of 37 / 52 / 67 KB against a 32 KB L1.5 I-cache):
# all four RK4 stages rolled (smaller code, explicit zeros in stage 1)
for _w in ("quadrotor", "quadbody", "quadbody64", "quadmrp64", "quadrotor64"):
    VARIANTS[_w] = {"base": {}, "roll2": dict(RDB_TUNE_ROLL=2)}


# instruction-cache experiment: the same thread layout, register pressure and amount of work per role, but every role runs the SAME code stream
# (results are wrong by construction — timing only): if the mean of the single-stream variants is well below the default, the gap is instruction fetch
def _same(mask, k, **kw):
    d = dict(RDB_TUNE_ROLL=1, **kw)
    for i in range(k):
        d[f"RDB_TUNE_C{i}"] = mask
    return d
VARIANTS["quadbody64"] = {"base": {}, "A4": _same("0x3Fu", 4, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1), "B4": _same("0x7C0u", 4, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1),
                          "C4": _same("0x3800u", 4, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1), "D4": _same("0x1C000u", 4, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1)}
VARIANTS["quadbody"] = {"base": {}, "A2": _same("0xFFFu", 2, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1), "B2": _same("0x1F000u", 2, RDB_TUNE_TILE=128, RDB_TUNE_MINB=1)}


# role balance (the single-stream experiment gave per-role times A 161 / B 184 / C 267 / D 274 us for the default fp64 split {0-5}{6-10}{11-13}{14-16}:
# u and w columns are the heavy ones, v columns nearly free): non-contiguous chunks that spread u and w
_T64 = dict(RDB_TUNE_ROLL=1, RDB_TUNE_TILE=64, RDB_TUNE_MINB=1)
VARIANTS["quadbody64"] = {"base": {}, "bal1": dict(_T64, RDB_TUNE_C0="0x6387u", RDB_TUNE_C1="0x18000u", RDB_TUNE_C2="0xC08u", RDB_TUNE_C3="0x1070u"),
                          "bal1c": dict(_T64, RDB_TUNE_C0="0x6187u", RDB_TUNE_C1="0x18200u", RDB_TUNE_C2="0xC08u", RDB_TUNE_C3="0x1070u")}
VARIANTS["quadrotor64"] = {"base": {}, "bal1": dict(_T64, RDB_TUNE_C0="0x6387u", RDB_TUNE_C1="0x18000u", RDB_TUNE_C2="0xC08u", RDB_TUNE_C3="0x1070u")}
VARIANTS["quadmrp64"] = {"base": {}, "balm": dict(_T64, RDB_TUNE_C0="0x23Fu", RDB_TUNE_C1="0xDC0u", RDB_TUNE_C2="0x3000u", RDB_TUNE_C3="0xC000u")}


# tile size between the two configurations in the library (wide 128 / narrow 32) for the fp32 rigid kernels: 64 knots, 2 CTAs per SM
for _w in ("quadrotor", "quadbody", "quadmrp", "bodyquat"):
    VARIANTS[_w] = {"t128": dict(RDB_TUNE_TILE=128, RDB_TUNE_MINB=1), "t64": dict(RDB_TUNE_TILE=64, RDB_TUNE_MINB=1), "t32": dict(RDB_TUNE_TILE=32, RDB_TUNE_MINB=1)}


VARIANTS["cartpole"].update({"t32_minb12": dict(RDB_TUNE_TILE=32, RDB_TUNE_MINB=12), "t64_minb6": dict(RDB_TUNE_TILE=64, RDB_TUNE_MINB=6)})


def build_variants(workload):
    import build as B
    os.makedirs(OUT, exist_ok=True)
    B.build()   # make sure abi.o / lie.o exist
    unit, defs = UNIT[workload]
    # stub object: every other unit symbol returns RDB_ERR_NOT_IMPLEMENTED
    stub_src = os.path.join(OUT, f"stubs_{workload}.cu")
    with open(stub_src, "w") as f:
        f.write('namespace rdb { struct KnotRequest; }\n')
        for name, _ in B.units():
            if name != unit:
                f.write(f'extern "C" int rdb_unit_{name}(const rdb::KnotRequest*) {{ return -2; }}\n')
    stub_obj = stub_src.replace(".cu", ".o")
    subprocess.check_call([B.NVCC] + B.ARCH + ["-Xcompiler", "-fPIC", "-c", "-o", stub_obj, stub_src])

    def one(tag, extra):
        obj = os.path.join(OUT, f"{workload}_{tag}.o")
        lib = os.path.join(OUT, f"{workload}_{tag}.so")
        flags = [f"-D{k}={v}" for k, v in {**defs, **extra}.items()] + [f"-DRDB_UNIT_NAME={unit}"]
        p = subprocess.run([B.NVCC] + B.ARCH + B.COMMON + ["-Xptxas", "-v"] + flags + ["-c", "-o", obj, os.path.join(CSRC, "unit.cu")],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            return tag, "COMPILE FAILED: " + p.stdout[-1500:]
        # resource line of the RK4 (or RK2 for satellite) Jacobian kernel
        want = "Li1E" if workload.startswith("satellite") else "Li3E"
        lines = p.stdout.splitlines()
        info = ""
        for i, l in enumerate(lines):
            if "Compiling entry function" in l and "knot_kernel" in l and want in l and "ELb1E" in l:
                info = " ".join(x.strip() for x in lines[i + 1:i + 4] if "registers" in x or "spill" in x)
        subprocess.check_call([B.NVCC] + B.ARCH + ["-shared", "-o", lib, obj, stub_obj, os.path.join(B.OBJ, "abi.o"), os.path.join(B.OBJ, "lie.o"),
                               os.path.join(B.OBJ, "custom.o"), os.path.join(B.OBJ, "layout.o"), "-cudart", "static", "-lnvrtc"])
        os.remove(obj)
        return tag, info

    only = os.environ.get("TUNE_ONLY")
    items = [(k, v) for k, v in VARIANTS[workload].items() if not only or k in only.split(",")]
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        for tag, info in ex.map(lambda kv: one(*kv), items):
            print(f"{workload}_{tag}: {info}", flush=True)


_RUN_ONE = r"""
import sys, os, json
sys.path.insert(0, sys.argv[1])
import numpy as np, torch
import rdb200 as rd
import bench
from oracle import rd_oracle as o
name = sys.argv[2]
wl = {"satellite32": "satellite", "quadrotor64": "quadrotor", "quadbody": "quadrotor", "quaderr": "quadrotor", "quadmrp": "quadrotor", "bodyquat": "quadrotor",
      "bodybody": "quadrotor", "bodybody64": "quadrotor", "quadmrpbody": "quadrotor", "quadlat": "quadrotor",
      "quadbody64": "quadrotor", "quadmrp64": "quadrotor"}.get(name, name)
desc, n, m, N, dtn, dt = bench.WORKLOADS[wl]
if name == "satellite32": dtn = "float32"
if name in ("quadrotor64", "quadbody64", "quadmrp64", "bodybody64"): dtn = "float64"
if len(sys.argv) > 3: N = int(sys.argv[3])
mk, Q = bench.gpu_model(wl, rd)
if name in ('quadbody', 'quadbody64'): mk = lambda: rd.Quadrotor(bodyframe=True)
if name in ('quadmrp', 'quadmrp64'): mk = lambda: rd.Quadrotor(rd.MRP)
if name == 'bodyquat': mk = lambda: rd.Body()
if name in ('bodybody', 'bodybody64'): mk = lambda: rd.Body(bodyframe=True)
if name == 'quadmrpbody': mk = lambda: rd.Quadrotor(rd.MRP, bodyframe=True)
model = mk(); h = model._h
n, m = h.n, h.m
nsets = 4
Zs = [torch.from_numpy(bench.make_inputs(n, m, N, dtn, i)).cuda() for i in range(nsets)]
ERR = name == "quaderr"
call = h.discrete_error_jacobian if ERR else h.discrete_jacobian
Js = [torch.empty((N, h.nerr + m, h.nerr) if ERR else (N, n + m, n), dtype=Zs[0].dtype, device="cuda") for _ in range(nsets)]
for i in range(5): call(Q.code, Zs[i % nsets], dt, J=Js[i % nsets])
torch.cuda.synchronize()
steps = 100
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps): call(Q.code, Zs[i % nsets], dt, J=Js[i % nsets])
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / steps * 1e3
omk, oQ = bench.oracle_model(wl)
if name in ('quadbody', 'quadbody64'): omk = lambda: o.quadrotor(o.ROT_QUAT, o.BODYFRAME)
if name in ('quadmrp', 'quadmrp64'): omk = lambda: o.quadrotor(o.ROT_MRP)
if name == 'bodyquat': omk = lambda: o.body()
if name in ('bodybody', 'bodybody64'): omk = lambda: o.body(o.ROT_QUAT, o.BODYFRAME)
if name == 'quadmrpbody': omk = lambda: o.quadrotor(o.ROT_MRP, o.BODYFRAME)
idx = np.arange(0, N, 4099)
if ERR:
    err = 0.0
else:
    ref = o.discrete_jacobian(omk(), oQ, Zs[0].cpu().numpy()[idx].astype(np.float64), dt)
    err = float(np.abs(Js[0].cpu().numpy()[idx] - ref).max())
es = Zs[0].element_size()
gbs = N * es * ((n + m) + (h.nerr * (h.nerr + m) if ERR else n * (n + m))) / (us * 1e-6) / 1e9
print(json.dumps({"us": us, "evals_per_s": N / (us * 1e-6), "GBs": gbs, "frac": gbs / 6551.4, "err": err}))
"""


_RUN_VALUE = r"""
import sys, os, json
sys.path.insert(0, sys.argv[1])
import numpy as np, torch
import rdb200 as rd
import bench
name = sys.argv[2]
wl = "cartpole" if name.startswith("cartpole") else "quadrotor"
desc, n, m, N, dtn, dt = bench.WORKLOADS[wl]
N = 1 << 22 if wl == "cartpole" else 1 << 21
mk, Q = bench.gpu_model(wl, rd)
h = mk()._h
nsets = 4
Zs = [torch.from_numpy(bench.make_inputs(n, m, N, dtn, i)).cuda() for i in range(nsets)]
Os = [torch.empty((N, n), dtype=Zs[0].dtype, device="cuda") for _ in range(nsets)]
out = {}
for label, op in (("discrete_dynamics", rd._abi.OP_DISCRETE_DYNAMICS), ("dynamics", rd._abi.OP_DYNAMICS)):
    plans = [rd._abi.Plan(h, op, Q.code, Z, dt, out=O) for Z, O in zip(Zs, Os)]
    for i in range(5): plans[i % nsets].launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2000000); e0.record()
    for i in range(50): plans[i % nsets].launch()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    es = Zs[0].element_size()
    out[label] = {"us": round(us, 1), "frac": round(N * es * (2 * n + m) / (us * 1e-6) / 1e9 / 6551.4, 3)}
print(json.dumps(out))
"""


def run_variants(workload, extra):
    libs = sorted(f for f in os.listdir(OUT) if f.startswith(workload + "_") and f.endswith(".so"))
    for lib in libs:
        env = dict(os.environ, RDB200_LIB=os.path.join(OUT, lib))
        p = subprocess.run([sys.executable, "-c", _RUN_VALUE if workload.endswith("_value") else _RUN_ONE, ROOT, workload] + extra, capture_output=True, text=True, env=env, timeout=600)
        last = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-300:]
        print(f"{lib:40s} {last}", flush=True)


if __name__ == "__main__":
    cmd, workload = sys.argv[1], sys.argv[2]
    if cmd == "build":
        build_variants(workload)
    else:
        run_variants(workload, sys.argv[3:])
