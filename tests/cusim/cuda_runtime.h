// cuda_runtime.h -- host-side half of the CPU SIMT emulator (TEST INFRASTRUCTURE ONLY, see cusim_device.h).
//
// Just enough of the CUDA runtime API for csrc/nd_b200.cu: "device" memory is host memory, streams run synchronously at
// the call, events are wall-clock stamps, stream capture records the launches into a replayable list, CUDA IPC handles are
// raw pointers (the emulated ranks of a multi-GPU test are threads of one process), run-time compiled kernels are shared
// objects built with g++ (nvrtc.h).
#pragma once
#include "cusim_device.h"

#include <functional>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInvalidConfiguration = 9, cudaErrorLaunchFailure = 719 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1 };

struct cusimStream;
struct cusimEvent;
struct cusimGraph;
struct cusimLibrary;
typedef cusimStream* cudaStream_t;
typedef cusimEvent* cudaEvent_t;
typedef cusimGraph* cudaGraph_t;
typedef cusimGraph* cudaGraphExec_t;
typedef cusimLibrary* cudaLibrary_t;
typedef void* cudaKernel_t;
struct cudaIpcMemHandle_t { char reserved[64]; };

namespace cusim {
void launch(dim3 grid, dim3 block, cudaStream_t st, std::function<void()> body);
}

const char* cudaGetErrorString(cudaError_t);
cudaError_t cudaGetLastError();
cudaError_t cudaSetDevice(int);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned flags);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemset(void* p, int v, size_t n);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t*, unsigned);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode);
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long);
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t);
cudaError_t cudaGraphDestroy(cudaGraph_t);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t);
cudaError_t cudaEventCreate(cudaEvent_t*);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t*, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*);
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned);
cudaError_t cudaIpcCloseMemHandle(void*);
cudaError_t cudaLibraryLoadData(cudaLibrary_t*, const void* code, void*, void*, unsigned, void*, void*, unsigned);
cudaError_t cudaLibraryUnload(cudaLibrary_t);
cudaError_t cudaLibraryGetKernel(cudaKernel_t*, cudaLibrary_t, const char* name);
cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t);
