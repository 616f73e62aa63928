// Hand-written build configuration for compiling the reference's vendored SEAL 4.1.1
// (/root/reference/depends/SEAL) WITHOUT its cmake build system. It states the option
// values a default Release configure with -DSEAL_USE_MSGSL=OFF -DSEAL_USE_ZLIB=OFF
// -DSEAL_USE_ZSTD=OFF -DSEAL_USE_INTEL_HEXL=OFF produces on x86-64 / gcc 13.
// TEST INFRASTRUCTURE ONLY (oracle/_ref); not part of the product.
#pragma once

#define SEAL_VERSION "4.1.1"
#define SEAL_VERSION_MAJOR 4
#define SEAL_VERSION_MINOR 1
#define SEAL_VERSION_PATCH 1

// C++17 features
#define SEAL_USE_STD_BYTE
#define SEAL_USE_ALIGNED_ALLOC
#define SEAL_USE_SHARED_MUTEX
#define SEAL_USE_IF_CONSTEXPR
#define SEAL_USE_MAYBE_UNUSED
#define SEAL_USE_NODISCARD
#define SEAL_USE_STD_FOR_EACH_N

// Security
#define SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT
#define SEAL_DEFAULT_PRNG Blake2xb

// Intrinsics (gcc, x86-64)
#define SEAL_USE_INTRIN
#define SEAL_USE___BUILTIN_CLZLL
#define SEAL_USE___INT128
#define SEAL_USE__ADDCARRY_U64
#define SEAL_USE__SUBBORROW_U64

// Zero memory functions
#define SEAL_USE_EXPLICIT_BZERO
