// units.h — the per-model compilation units (unit.cu, built many times by build.py) as seen from abi.cu.
#pragma once
#include "launch.cuh"

#define RDB_DECL_UNIT(name) extern "C" int rdb_unit_##name(const rdb::KnotRequest*);
#define RDB_DECL_BOTH(name) RDB_DECL_UNIT(name##_f32) RDB_DECL_UNIT(name##_f64)
#define RDB_DECL_RIGID(kind) \
    RDB_DECL_BOTH(kind##_quat_world) RDB_DECL_BOTH(kind##_quat_body) RDB_DECL_BOTH(kind##_mrp_world) \
    RDB_DECL_BOTH(kind##_mrp_body) RDB_DECL_BOTH(kind##_rp_world) RDB_DECL_BOTH(kind##_rp_body)
RDB_DECL_BOTH(cartpole)
RDB_DECL_BOTH(di1) RDB_DECL_BOTH(di2) RDB_DECL_BOTH(di3)
RDB_DECL_RIGID(quad)
RDB_DECL_RIGID(body)

namespace rdb {
using UnitFn = int (*)(const KnotRequest*);

// kind/rot/frame as in models.cuh; D only for double integrators; dtype 0 = f32, 1 = f64
inline UnitFn find_unit(int kind, int rot, int frame, int D, int dtype) {
#define RDB_PICK(name) (dtype == 0 ? rdb_unit_##name##_f32 : rdb_unit_##name##_f64)
#define RDB_PICK_RIGID(kind_)                                                                      \
    switch (rot * 2 + frame) {                                                                     \
        case 2: return RDB_PICK(kind_##_quat_world); case 3: return RDB_PICK(kind_##_quat_body);   \
        case 4: return RDB_PICK(kind_##_mrp_world);  case 5: return RDB_PICK(kind_##_mrp_body);    \
        case 6: return RDB_PICK(kind_##_rp_world);   case 7: return RDB_PICK(kind_##_rp_body);     \
        default: return nullptr;                                                                   \
    }
    if (dtype != 0 && dtype != 1) return nullptr;
    switch (kind) {
        case KIND_CARTPOLE: return RDB_PICK(cartpole);
        case KIND_DOUBLE_INTEGRATOR:
            return D == 1 ? RDB_PICK(di1) : D == 2 ? RDB_PICK(di2) : D == 3 ? RDB_PICK(di3) : nullptr;
        case KIND_QUADROTOR: RDB_PICK_RIGID(quad)
        case KIND_BODY: RDB_PICK_RIGID(body)
    }
    return nullptr;
#undef RDB_PICK
#undef RDB_PICK_RIGID
}
}  // namespace rdb
