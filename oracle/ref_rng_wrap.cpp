// ref_rng_wrap.cpp -- compiles the REFERENCE's own optix/random.hpp (from where it lies under /root/reference,
// never copied) into oracle/_ref/librefrng.so so that the oracle's tea4/lcg/rnd restatement can be checked
// against the reference implementation itself. Test infrastructure only; built only when /root/reference exists.
#include <cstdint>
#define __host__
#define __device__
#define __inline__ inline
#include REF_RANDOM_HPP

extern "C" {
uint32_t ref_tea4(uint32_t a, uint32_t b) { return tea<4>(a, b); }
void ref_rnd_sequence(uint32_t seed, int n, float* out) { for (int i = 0; i < n; i++) out[i] = rnd(seed); }
uint32_t ref_lcg(uint32_t* prev) { return lcg(*prev); }
}
