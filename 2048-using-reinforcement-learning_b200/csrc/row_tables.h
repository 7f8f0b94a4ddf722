// row_tables.h -- host-side construction of the two 65,536-entry row tables.
// Plain C++ (no CUDA) so that tests/host_emul can build the very same tables.
#pragma once
#include <stdint.h>

namespace g2048 {

// Row semantics of env:116-168 == agent:213-242: drop the zeros, scan left to right merging an
// equal neighbour once, pad with zeros.
//   row[r]  = the 4 result nibbles (a merged exponent of 16 saturates to 15)
//   code[r] = merges of that move, one nibble each (first merge in the low nibble):
//             0 = none, else merged exponent - 1, i.e. the merge scores 2 << nibble.
inline void build_row_tables(uint16_t *row, uint8_t *code)
{
    for (int r = 0; r < 65536; ++r) {
        int t[4], m = 0;
        for (int j = 0; j < 4; ++j) { int e = (r >> (4 * j)) & 15; if (e) t[m++] = e; }
        int out[4] = {0, 0, 0, 0}, k = 0, nmerge = 0, c = 0;
        for (int j = 0; j < m;) {
            if (j + 1 < m && t[j] == t[j + 1]) {
                int e = t[j] + 1;                     // merged exponent, 2..16
                c |= (e - 1) << (4 * nmerge++);
                out[k++] = e > 15 ? 15 : e;           // nibble saturation, visible in the code (e - 1 == 15)
                j += 2;
            } else {
                out[k++] = t[j];
                j += 1;
            }
        }
        row[r] = (uint16_t)(out[0] | (out[1] << 4) | (out[2] << 8) | (out[3] << 12));
        code[r] = (uint8_t)c;
    }
}

}  // namespace g2048
