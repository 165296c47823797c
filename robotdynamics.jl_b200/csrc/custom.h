// custom.h — host entry points of custom.cu (NVRTC-compiled user models).
#pragma once
#include "launch.cuh"

namespace rdb {
struct CustomModel;
int custom_check(int n, int m, const char* body, int nparams, int dtype, int rot = 0, int frame = 0);   // compile only; 0 = ok (no GPU needed)
const char* custom_last_log();                                               // NVRTC log of this thread's last compile
// rot != 0: RigidBody{R} with a user wrench body (mp carries mass and inertia); rot == 0: Euclidean model with a user f body
CustomModel* custom_create(int n, int m, const char* body, const double* params, int nparams, int rot = 0, int frame = 0,
                            const ModelParams<double>* mp = nullptr);
void custom_destroy(CustomModel* c);
int custom_run(CustomModel* c, const KnotRequest& r);
}  // namespace rdb
