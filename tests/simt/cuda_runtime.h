// Host stand-in for <cuda_runtime.h> in the SIMT-emulator build (tests/simt): everything lives in simt.h.
#pragma once
#include "simt.h"
