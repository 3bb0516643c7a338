// nvrtc.h -- run-time compilation for the CPU SIMT emulator (TEST INFRASTRUCTURE ONLY, see cusim_device.h).
// "Compiling" a program = g++ -shared of the generated source against cusim_device.h; the "cubin" is the path of the
// resulting shared object, which cudaLibraryLoadData dlopens.
#pragma once
#include <stddef.h>

typedef int nvrtcResult;
enum { NVRTC_SUCCESS = 0, NVRTC_ERROR_COMPILATION = 6, NVRTC_ERROR_INVALID_INPUT = 3 };
struct cusimProgram;
typedef cusimProgram* nvrtcProgram;

nvrtcResult nvrtcCreateProgram(nvrtcProgram*, const char* src, const char* name, int nh, const char* const* headers, const char* const* names);
nvrtcResult nvrtcDestroyProgram(nvrtcProgram*);
nvrtcResult nvrtcAddNameExpression(nvrtcProgram, const char* expr);
nvrtcResult nvrtcCompileProgram(nvrtcProgram, int nopts, const char* const* opts);
nvrtcResult nvrtcGetProgramLogSize(nvrtcProgram, size_t*);
nvrtcResult nvrtcGetProgramLog(nvrtcProgram, char*);
nvrtcResult nvrtcGetCUBINSize(nvrtcProgram, size_t*);
nvrtcResult nvrtcGetCUBIN(nvrtcProgram, char*);
nvrtcResult nvrtcGetLoweredName(nvrtcProgram, const char* expr, const char** lowered);
