/*
 * eaof_sweep.h — C ABI of the one exchange step of the ORB path: the cross-frame brute-force Hamming sweep over
 * frames that were extracted on different GPUs (BASELINE.json configs[4]).
 *
 * The reference has no distributed layer at all; what this replaces is the loop a caller would write
 * around ORBmatcher::SearchByBoW (src/ORBmatcher.cc:159-288) with one vocabulary node holding every feature: for each
 * (query frame, target frame) pair of a global list, the matches of the pair.  One process per GPU; every rank holds
 * the descriptor blocks of the frames it extracted (block = one frame: `block_stride` rows of 32 bytes, keypoint
 * angles, a feature count).  The blocks are all-gathered ONCE with ncclAllGather over NVLink/NVSwitch, the pair list
 * is partitioned by the caller (round-robin: eaof/shard.py sweep_pairs), and each rank matches its share with the
 * kernels behind eaof_match_bruteforce_batch_device.  There is no other collective on the data path.
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already loaded — e.g. PyTorch's — or the
 * system one), so libeaof_orb.so itself loads on machines without NCCL; every NCCL failure is reported as
 * EAOF_ERR_NCCL with ncclGetErrorString / ncclGetLastError in eaof_last_error().  world == 1 needs no NCCL.
 *
 * Same conventions as eaof_orb.h: 0 / negative EAOF_ERR_*, eaof_last_error(), no CPU fallback.
 */
#ifndef EAOF_SWEEP_H
#define EAOF_SWEEP_H

#include <stddef.h>
#include <stdint.h>

#include "eaof_match.h"

#ifdef __cplusplus
extern "C" {
#endif

#define EAOF_SWEEP_ID_BYTES 128 /* sizeof(ncclUniqueId) */

typedef struct eaof_sweep eaof_sweep;

/* ncclGetUniqueId: called by one rank; the 128 bytes travel to the other ranks by whatever the host program uses
 * (torch.distributed broadcast in bench.py, a file or a socket in a C++ program). */
int eaof_sweep_unique_id(uint8_t id[EAOF_SWEEP_ID_BYTES]);
/* NCCL version the process bound (major*10000 + minor*100 + patch), for logs. */
int eaof_sweep_nccl_version(int* version);

/* ncclCommInitRank on `device` (collective: every rank of the world calls it with the same id).  world == 1 creates a
 * handle without a communicator (id may be NULL). */
int eaof_sweep_create(const uint8_t id[EAOF_SWEEP_ID_BYTES], int rank, int world, int device, eaof_sweep** out);
void eaof_sweep_destroy(eaof_sweep* s);

/* The exchange step alone: every rank contributes `blocks_per_rank` blocks (ranks with fewer frames pad with blocks
 * of count 0); on return (asynchronously, on `stream`) d_*_all hold world*blocks_per_rank blocks, rank r's at
 * [r*blocks_per_rank, (r+1)*blocks_per_rank).  One ncclGroup of three ncclAllGather (descriptors, angles, counts).
 * stream: a cudaStream_t (NULL = the legacy default stream). */
int eaof_sweep_allgather_blocks(eaof_sweep* s, int blocks_per_rank, int block_stride, const uint8_t* d_desc_local,
                                const float* d_angle_local, const int* d_count_local, uint8_t* d_desc_all,
                                float* d_angle_all, int* d_count_all, void* stream);

/* Exchange + matching of this rank's pairs in one call, all on the matcher's stream (asynchronous): the all-gather
 * above, then eaof_match_bruteforce_batch_device semantics over pairs (pair_q[p], pair_t[p]) given as indices into
 * the gathered block array, in chunks of the matcher's max_pairs.  Outputs laid out [pair][block_stride] like
 * eaof_match_bow (mode EAOF_BOW_KF_FRAME / EAOF_BOW_KF_KF).  pair_q / pair_t are host arrays. */
int eaof_sweep_match(eaof_sweep* s, eaof_matcher* m, int mode, float nnratio, int check_orientation, int blocks_per_rank,
                     int block_stride, const uint8_t* d_desc_local, const float* d_angle_local, const int* d_count_local,
                     uint8_t* d_desc_all, float* d_angle_all, int* d_count_all, int n_pairs, const int* pair_q,
                     const int* pair_t, int* d_match, int* d_dist, int* d_nmatches);

/* Device time of the last all-gather in milliseconds (CUDA events on the stream it ran on; call after that stream
 * was synchronised) and the bytes this rank received in it. */
int eaof_sweep_last_allgather(eaof_sweep* s, float* ms, long long* bytes_received);

int eaof_sweep_rank(const eaof_sweep* s);
int eaof_sweep_world(const eaof_sweep* s);

#ifdef __cplusplus
}
#endif
#endif /* EAOF_SWEEP_H */
