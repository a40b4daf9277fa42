// declaration-only stand-in for the reference's logging macros (src/debug.h:121-157)
#ifndef PBA_PROOF_DEBUG_H
#define PBA_PROOF_DEBUG_H
#include <cstdio>
#define Info(...) do { std::printf(__VA_ARGS__); } while (0)
#define Warn(...) do { std::fprintf(stderr, __VA_ARGS__); } while (0)
#endif
