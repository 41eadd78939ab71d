/* TEST INFRASTRUCTURE ONLY: empty stand-in so the reference .cu kernel text compiles on the host (see cuda_on_cpu.h). */
#pragma once
#include "cuda_on_cpu.h"
