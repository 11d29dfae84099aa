// Host build of csrc/tg_rng.cuh (the device RNG) for the CPU test: reads an MT19937 state (624 words + pos) and a draw program
// from stdin, prints the draws.  __device__ qualifiers are defined away; the code under test is the header itself.
#define __device__
#define __forceinline__ inline
#define __noinline__
#include <cstdio>
#include <vector>
#include "../../tactile_gym_b200/csrc/tg_rng.cuh"

int main()
{
    std::vector<uint32_t> key(MT_N);
    int pos, nprog, rounds;
    for (auto& k : key) if (scanf("%u", &k) != 1) return 1;
    if (scanf("%d %d %d", &pos, &nprog, &rounds) != 3) return 1;
    std::vector<int> kind(nprog);
    std::vector<double> lo(nprog), hi(nprog);
    for (int i = 0; i < nprog; i++) if (scanf("%d %lf %lf", &kind[i], &lo[i], &hi[i]) != 3) return 1;
    MtState s{key.data(), pos};
    for (int r = 0; r < rounds; r++)
        for (int i = 0; i < nprog; i++) printf("%.17g\n", mt_draw(s, kind[i], lo[i], hi[i], -7.0));
    printf("%d\n", s.pos);
    return 0;
}
