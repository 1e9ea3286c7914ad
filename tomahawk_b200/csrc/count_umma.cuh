// count_umma.cuh -- tensor-core (tcgen05 / TMEM) variant of the count kernel.
// Placeholder interface: filled in once the LOP3+POPC path is parity-green.
#pragma once
#include <string>
#include "common.cuh"
#include "count_popc.cuh"

namespace twkb {

constexpr uint32_t UMMA_TILE_M = 128;
constexpr uint32_t UMMA_TILE_N = 128;

struct UmmaOperand {
    bool valid = false;
    uint32_t Kbytes = 0;
};

inline bool umma_supported() { return false; }
inline int umma_prepare(UmmaOperand&, const uint32_t*, uint32_t, uint32_t, uint32_t, cudaStream_t, std::string&) { return 0; }
inline cudaError_t umma_launch(UmmaOperand&, const CountArgs&, const DevParams&, uint32_t, cudaStream_t) { return cudaErrorNotSupported; }
inline void umma_release(UmmaOperand&) {}

}  // namespace twkb
