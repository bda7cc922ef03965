// tally_kernel.cuh -- device merge (Donatello semantics) + per-read tally. (filled in below)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace elector {}
