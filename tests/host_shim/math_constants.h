// Test infrastructure (see cuda_runtime.h in this directory): the two constants the arithmetic headers take from CUDA's math_constants.h
#pragma once
#define CUDART_INF __longlong_as_double(0x7ff0000000000000LL)
#define CUDART_NAN __longlong_as_double((long long)0xfff8000000000000ULL)
