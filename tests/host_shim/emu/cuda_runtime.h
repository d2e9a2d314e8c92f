// TEST INFRASTRUCTURE.  The slice of the CUDA runtime API that libspparks_cuda uses, implemented on
// the host by cuda_emu.cpp (see cuda_emu.h).  Streams are synchronous.
#pragma once
#include "cuda_emu.h"

typedef int cudaError_t;
enum : int { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600,
             cudaErrorNotSupported = 801, cudaErrorUnknown = 999 };
typedef struct emuStream_st *cudaStream_t;
typedef struct emuEvent_st *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaEventDefault = 0, cudaEventDisableTiming = 2 };
enum { cudaHostRegisterDefault = 0 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaEnableDefault = 0 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int dev);
cudaError_t emuMalloc(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return emuMalloc((void **)p, bytes); }
cudaError_t cudaFree(void *p);
cudaError_t emuMallocHost(void **p, size_t bytes);
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return emuMallocHost((void **)p, bytes); }
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaHostRegister(void *p, size_t bytes, unsigned flags);
cudaError_t cudaHostUnregister(void *p);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s = nullptr);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemsetAsync(void *p, int value, size_t bytes, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void *p, int value, size_t bytes);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventQuery(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
cudaError_t cudaGetDriverEntryPoint(const char *symbol, void **fn, unsigned long long flags,
                                    cudaDriverEntryPointQueryResult *status = nullptr);
cudaError_t emuFuncSetAttribute(int attr, int value);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute attr, int value) {
  return emuFuncSetAttribute((int)attr, value);
}
