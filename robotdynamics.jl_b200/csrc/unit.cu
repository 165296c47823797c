// unit.cu — one compilation unit per (model family, rotation, frame, dtype): instantiates knot_kernel and
// rollout_kernel for every quadrature rule.  Compiled many times with different -D flags (see build.py):
//   -DRDB_UNIT_NAME=<symbol suffix> -DRDB_KIND=<0..3> -DRDB_ROT=<0..3> -DRDB_FRAME=<0|1> -DRDB_DI_D=<1..3> -DRDB_DTYPE=<0|1>
#include "launch.cuh"

#if !defined(RDB_UNIT_NAME) || !defined(RDB_KIND) || !defined(RDB_DTYPE)
#error "unit.cu needs -DRDB_UNIT_NAME, -DRDB_KIND, -DRDB_DTYPE"
#endif
#ifndef RDB_ROT
#define RDB_ROT 0
#endif
#ifndef RDB_FRAME
#define RDB_FRAME 0
#endif
#ifndef RDB_DI_D
#define RDB_DI_D 1
#endif

namespace {
#if RDB_KIND == 0
template <class T> using UnitModel = rdb::Cartpole<T>;
#elif RDB_KIND == 3
template <class T> using UnitModel = rdb::DoubleIntegrator<T, RDB_DI_D>;
#else
template <class T> using UnitModel = rdb::RigidBody<T, RDB_KIND, RDB_ROT, RDB_FRAME>;
#endif
}  // namespace

#define RDB_CAT2(a, b) a##b
#define RDB_CAT(a, b) RDB_CAT2(a, b)

extern "C" int RDB_CAT(rdb_unit_, RDB_UNIT_NAME)(const rdb::KnotRequest* r) {
#if RDB_DTYPE == 0
    return rdb::run_t<UnitModel, float>(*r);
#else
    return rdb::run_t<UnitModel, double>(*r);
#endif
}
