/* Minimal C client of the C ABI (include/g2048.h): no CUDA headers, no C++ -- the boundary a
 * non-Python host binds.  Plays 4 environments for 100 random steps through the host-buffer entry
 * points and prints the scores.  Build:  gcc examples/abi_client.c -Iinclude -L<dir of libg2048.so>
 * -lg2048 -o abi_client     (on a machine without a GPU it reports G2048_ENODEVICE and exits 3). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "g2048.h"

int main(void)
{
    enum { N = 4, STEPS = 100 };
    uint64_t boards[N] = {0};
    int32_t score[N] = {0};
    uint8_t highest[N] = {0}, actions[N], valid[N], done[N];
    uint32_t spawn_ctr[N] = {0};
    double reward[N];

    printf("libg2048 ABI version %d\n", g2048_abi_version());
    int rc = g2048_init(0);
    if (rc != G2048_OK) {
        printf("g2048_init: %d (%s)\n", rc, g2048_last_error());
        return rc == G2048_ENODEVICE ? 3 : 1;
    }
    if ((rc = g2048_host_env_reset(boards, score, highest, spawn_ctr, N, 2048u, 0u)) != G2048_OK) goto fail;
    srand(1);
    for (int t = 0; t < STEPS; ++t) {
        for (int i = 0; i < N; ++i) actions[i] = (uint8_t)(rand() & 3);
        rc = g2048_host_env_step(boards, actions, NULL, score, highest, spawn_ctr, reward, NULL, valid, NULL, done, N, 2048u, 0u);
        if (rc != G2048_OK) goto fail;
    }
    for (int i = 0; i < N; ++i)
        printf("env %d: board %016llx score %d highest tile %d\n", i, (unsigned long long)boards[i], score[i],
               highest[i] ? 1 << highest[i] : 0);
    return 0;
fail:
    printf("error %d: %s\n", rc, g2048_last_error());
    return 1;
}
