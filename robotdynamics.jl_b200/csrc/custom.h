// custom.h — host entry points of custom.cu (NVRTC-compiled user models).
#pragma once
#include "launch.cuh"

namespace rdb {
struct CustomModel;
int custom_check(int n, int m, const char* body, int nparams, int dtype);   // compile only; 0 = ok (no GPU needed)
const char* custom_last_log();                                               // NVRTC log of this thread's last compile
CustomModel* custom_create(int n, int m, const char* body, const double* params, int nparams);
void custom_destroy(CustomModel* c);
int custom_run(CustomModel* c, const KnotRequest& r);
}  // namespace rdb
