/*
 * hb_codebook.c -- host-side Huffman codebook over 256 byte symbols (plain C, no device work).
 *
 * Drop-in for the reference's BuildTree + GenerateCodes (huffTree.h:55-94) followed by the
 * table flatten in loadData (load_data.h:40-47).  The reference leans on
 * std::priority_queue<INode*, std::vector<INode*>, NodeCmp> (huffTree.h:57), so equal-weight
 * ties are resolved by libstdc++'s binary heap.  To be bit-identical without linking libstdc++
 * the two heap primitives are restated here the way bits/stl_heap.h runs them:
 *     push  = append, then sift the value up while comp(parent, value)
 *     pop   = save the LAST element, move the root into its slot, sift the hole at the root
 *             down to a leaf always following the child that does not lose comp(), then sift
 *             the saved element up from that leaf
 * with comp(a, b) = weight[a] > weight[b] (NodeCmp, huffTree.h:50-53).
 *
 * Differences from the reference, all deliberate (SURVEY.md sections 8 a-3, 9):
 *   - weights are int64 (the reference's `const int f`, huffTree.h:22, wraps above INT_MAX);
 *   - an all-zero histogram returns all-zero tables (the reference calls top() on an empty queue);
 *   - codes longer than 31 bits are refused (the encode parity domain).
 */
#include <string.h>

#include "../../include/huffman_b200.h"

#define HB_MAX_NODES (2 * HB_NUM_SYMBOLS - 1)

typedef struct {
    long long weight[HB_MAX_NODES];
    short parent[HB_MAX_NODES];
    unsigned char is_right[HB_MAX_NODES];   /* 1 when this node is its parent's `right` child */
    short slots[HB_NUM_SYMBOLS];            /* the heap: node ids */
    int count;
} hb_forest;

static int heavier(const hb_forest *t, int a, int b) { return t->weight[a] > t->weight[b]; }

static void sift_up(hb_forest *t, int hole, int floor_index, int node)
{
    while (hole > floor_index) {
        int up = (hole - 1) / 2;
        if (!heavier(t, t->slots[up], node))
            break;
        t->slots[hole] = t->slots[up];
        hole = up;
    }
    t->slots[hole] = (short)node;
}

static void heap_insert(hb_forest *t, int node)
{
    t->slots[t->count] = (short)node;
    t->count++;
    sift_up(t, t->count - 1, 0, node);
}

static int heap_extract(hb_forest *t)
{
    const int root = t->slots[0];
    const int n = t->count - 1;             /* heap length after the extraction */
    if (n > 0) {
        const int carried = t->slots[n];
        int hole = 0, kid = 0;
        t->slots[n] = (short)root;
        while (kid < (n - 1) / 2) {
            kid = 2 * (kid + 1);            /* right child */
            if (heavier(t, t->slots[kid], t->slots[kid - 1]))
                kid--;                      /* right loses: take the left one */
            t->slots[hole] = t->slots[kid];
            hole = kid;
        }
        if ((n & 1) == 0 && kid == (n - 2) / 2) {   /* a last node with only a left child */
            kid = 2 * (kid + 1);
            t->slots[hole] = t->slots[kid - 1];
            hole = kid - 1;
        }
        sift_up(t, hole, 0, carried);
    }
    t->count = n;
    return root;
}

int hb_build_codebook(const uint64_t hist[256], uint32_t codewords[256],
                      uint32_t codewordlens[256])
{
    hb_forest t;
    short leaf_of[HB_NUM_SYMBOLS];
    int nodes = 0, longest = 0;

    if (!hist || !codewords || !codewordlens)
        return HB_ERR_ARG;
    memset(codewords, 0, HB_NUM_SYMBOLS * sizeof(uint32_t));
    memset(codewordlens, 0, HB_NUM_SYMBOLS * sizeof(uint32_t));
    t.count = 0;

    /* leaves enter the queue in symbol order (huffTree.h:59-63) */
    for (int s = 0; s < HB_NUM_SYMBOLS; s++) {
        leaf_of[s] = -1;
        if (hist[s] == 0)
            continue;
        t.weight[nodes] = (long long)hist[s];
        t.parent[nodes] = -1;
        t.is_right[nodes] = 0;
        leaf_of[s] = (short)nodes;
        heap_insert(&t, nodes);
        nodes++;
    }
    if (nodes == 0)
        return 0;

    /* huffTree.h:64-74: the first node popped becomes `left` (edge 0), the second `right` (edge 1) */
    while (t.count > 1) {
        const int first = heap_extract(&t);
        const int second = heap_extract(&t);
        t.weight[nodes] = t.weight[first] + t.weight[second];
        t.parent[nodes] = -1;
        t.is_right[nodes] = 0;
        t.parent[first] = (short)nodes;
        t.parent[second] = (short)nodes;
        t.is_right[second] = 1;
        heap_insert(&t, nodes);
        nodes++;
    }

    /* Walk each leaf up to the root.  The edge next to the root is the MSB of the codeword
     * (GenerateCodes pushes root-first, load_data.h:44-45 weights bit i with 2^(len-1-i)). */
    for (int s = 0; s < HB_NUM_SYMBOLS; s++) {
        unsigned long long bits = 0;
        int depth = 0;
        if (leaf_of[s] < 0)
            continue;
        for (int n = leaf_of[s]; t.parent[n] >= 0; n = t.parent[n]) {
            if (depth < 64)
                bits |= (unsigned long long)t.is_right[n] << depth;
            depth++;
        }
        if (depth > HB_MAX_CODE_LEN)
            return HB_ERR_CODELEN;
        codewords[s] = (uint32_t)bits;
        codewordlens[s] = (uint32_t)depth;
        if (depth > longest)
            longest = depth;
    }
    return longest;
}

uint64_t hb_bits_from_hist(const uint64_t hist[256], const uint32_t codewordlens[256])
{
    uint64_t bits = 0;
    for (int s = 0; s < HB_NUM_SYMBOLS; s++)
        bits += hist[s] * (uint64_t)codewordlens[s];
    return bits;
}

int hb_shard_offsets(const uint64_t *shard_bits, int n_shards, uint64_t *start_bits,
                     uint64_t *total_bits)
{
    uint64_t run = 0;
    if (!shard_bits || !start_bits || n_shards < 0)
        return HB_ERR_ARG;
    for (int r = 0; r < n_shards; r++) {
        start_bits[r] = run;
        run += shard_bits[r];
    }
    if (total_bits)
        *total_bits = run;
    return HB_OK;
}
