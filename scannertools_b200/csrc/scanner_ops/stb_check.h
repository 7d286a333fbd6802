// Shared helper for the Scanner kernel classes: turn a non-zero C-ABI status into the
// reference's failure mode (abort the worker, like CU_CHECK; optical_flow_kernel_gpu.cpp:97).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "stb.h"

#define STB_CHECK(call)                                                                    \
  do {                                                                                     \
    int stb_rc__ = (call);                                                                 \
    if (stb_rc__ != 0) {                                                                   \
      fprintf(stderr, "scannertools_b200: %s failed with status %d: %s (%s:%d)\n", #call,  \
              stb_rc__, stb_last_error(), __FILE__, __LINE__);                              \
      abort();                                                                             \
    }                                                                                      \
  } while (0)

// unrecoverable input (the reference uses glog LOG(FATAL), image_decoder_kernel_cpu.cpp:28): report and
// abort the worker -- never reinterpret a frame of the wrong type
#define STB_FATAL(msg)                                                              \
  do {                                                                              \
    fprintf(stderr, "scannertools_b200: %s (%s:%d)\n", (msg), __FILE__, __LINE__);  \
    abort();                                                                        \
  } while (0)
