/* TEST INFRASTRUCTURE ONLY (oracle). Philox-4x32-10 counter-based RNG, restated from the published
 * algorithm (Salmon, Moraes, Dror, Shaw: "Parallel Random Numbers: As Easy as 1, 2, 3", SC'11; Random123
 * philox4x32-10). The reference itself uses rand::thread_rng() (OS-seeded ChaCha12, irreproducible:
 * /root/reference/src/breakout-game/src/mechanics.rs:103,
 * /root/reference/src/ql-with-tensorflow/src/learn/self_driving_tf_q_learner.rs:106), so every random draw on
 * the hot path is an explicit input; this generator is OUR definition of those inputs, shared by spec (not by
 * code) with the CUDA product. Known-answer vectors from the Random123 distribution are checked in
 * oracle/selftest.c and tests/test_oracle.py. */
#ifndef QLC_ORACLE_PHILOX_H
#define QLC_ORACLE_PHILOX_H
#include <stdint.h>

static inline void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * (uint64_t)c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * (uint64_t)c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream tags (counter word 3) used by the spec, see DESIGN.md "Random inputs". */
#define ORC_STREAM_RESET  0x52455345u /* 'RESE': initial ball direction per (env, episode) */
#define ORC_STREAM_SAMPLE 0x53414D50u /* 'SAMP': minibatch index draws per (call, draw)   */
#define ORC_STREAM_ACTION 0x41435449u /* 'ACTI': synthetic random policy per (env, step)   */
#endif
