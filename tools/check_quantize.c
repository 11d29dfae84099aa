/* Exhaustive check behind tg_raster.cuh:quantize().
 * The reference computes uint8(((clip(pen, 0, 0.05) / 0.05) * 255)) in float32 (sensors/tactile_sensor.py:281-284).
 * The kernel replaces the IEEE division by  q0 = pen * 20;  q = fma(fma(-0.05f, q0, pen), 20, q0).
 * This program walks EVERY float in [0, 0.05f] (1 028 443 342 values) and counts the inputs whose final uint8
 * differs; it must print 0.   gcc -O2 -mfma -ffp-contract=off check_quantize.c -lm   [stride] samples every stride-th value */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char** argv)
{
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    const float c = 0.05f, r = 20.0f;
    uint32_t hi;
    memcpy(&hi, &c, 4);
    long bad = 0, n = 0;
    for (uint64_t u = 0; u <= hi; u += stride, n++) {
        uint32_t uu = (uint32_t)u;
        float a;
        memcpy(&a, &uu, 4);
        const unsigned char ref = (unsigned char)((a / c) * 255.0f);
        const float q0 = a * r;
        const float q = fmaf(fmaf(-c, q0, a), r, q0);
        const unsigned char fast = (unsigned char)(q * 255.0f);
        if (ref != fast) bad++;
    }
    printf("%ld values checked, %ld uint8 mismatches\n", n, bad);
    return bad != 0;
}
