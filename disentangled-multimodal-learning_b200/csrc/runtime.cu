// Runtime guard: the library is built for sm_100a only and refuses to run anywhere else.
#include "common.cuh"

extern "C" {

const char* dml_version(void) { return "dml_b200 0.1.0 (sm_100a)"; }

int dml_runtime_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return (int)e;
  return major == 10 ? DML_OK : DML_EUNSUPPORTED;
}

}  // extern "C"
