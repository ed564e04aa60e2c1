// stub for the CUDA emulator build (tests/emu/cuda_emu.h): TEST INFRASTRUCTURE
#pragma once
#include "../cuda_emu.h"
