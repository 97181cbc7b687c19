// Opaque stand-ins so that /root/reference/src/tetrahedra_tracer.cu (which uses no OptiX symbol,
// only includes tetrahedra_tracer.h) compiles without the OptiX SDK.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cuda.h>
typedef struct OptixDeviceContext_t *OptixDeviceContext;
typedef unsigned long long OptixTraversableHandle;
typedef struct OptixModule_t *OptixModule;
typedef struct OptixPipeline_t *OptixPipeline;
typedef struct OptixProgramGroup_t *OptixProgramGroup;
typedef struct OptixShaderBindingTable { CUdeviceptr a[16]; } OptixShaderBindingTable;
