// Force-included (-include) in front of the unmodified reference sources when building oracle/_ref.
// 1. the headers the reference relies on MSVC to pull in transitively, and every standard header it includes
//    AFTER defining its function-like `max` macro (opticalFlowCalc.h:14) — including them first makes the
//    later #include a no-op, so the macro can not mangle libstdc++;
// 2. the two Win32 spellings in opticalFlowCalc.h:12.
#pragma once
#include <math.h>
#include <stdio.h>
#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <exception>
#include <iostream>
#include <stdexcept>
#include <string>

#define __declspec(x)
#define __stdcall
extern "C" void OutputDebugStringA(const char*);
