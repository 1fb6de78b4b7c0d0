// Error reporting shared by the CUDA product build and the test-only host emulation (no CUDA dependency).
#pragma once
#include "../../include/mezzo_b200.h"

namespace mz {
void set_error(const char *fmt, ...);
const char *last_error();  // the calling thread's message (the buffer is thread-local)
}

#define MZ_REQUIRE(cond, code, ...)  \
    do {                             \
        if (!(cond)) {               \
            mz::set_error(__VA_ARGS__); \
            return (code);           \
        }                            \
    } while (0)
