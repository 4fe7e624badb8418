// Chunk tables ("plans") of the shared-memory-staged kernels (staged.cu).
#pragma once
#include <stdint.h>
#include <vector>

struct BlockChunk {       // a run of consecutive nodes of the block system
    int32_t n0, nn;                 // first node, number of nodes
    int32_t rp_off, rp_len;         // 16-byte aligned slice of the node row pointers
    int32_t cA_off, cA_len;         // aligned slice of the primary column indices
    int32_t cB_off, cB_len;         // aligned slice of the secondary column indices
    int64_t v_rel;                  // offset of the chunk inside every row group's value run
    int32_t v_len, pad;             // values per row group
};

struct SpmmChunk {        // a run of consecutive rows of a scalar CSR matrix
    int32_t r0, nrows;
    int32_t rp_off, rp_len;
    int32_t v_off, v_len;
    int32_t c_off, c_len;
};

// cfg = index of the pipeline configuration (threads / stages / stage capacity) the plan was cut for
struct BlockPlan { BlockChunk* chunks = nullptr; int nchunks = 0; int cfg = 0; };
struct SpmmPlan { SpmmChunk* chunks = nullptr; int nchunks = 0; int cfg = 0; };
