/* merkle.c — plonky2 MerkleTree<F, PoseidonHash>.  TEST INFRASTRUCTURE (see p2oracle.h).
 * Restates plonky2 0.2.2 hash/merkle_tree.rs (fill_digests_buf / fill_subtree / prove) and
 * hash/merkle_proofs.rs (verify_merkle_proof_to_cap) — not on disk; SURVEY A.5.
 * Pinned by the Merkle paths inside the ten stored proofs of qbench_data/example.bin
 * (leaf hashing of 135/20/16/32-wide leaves, sibling order, cap index = top bits). */
#include <stdlib.h>
#include <string.h>

#include "gl_inline.h"

/* digests_buf: 2*(n_leaves-1) digests for a subtree; layout
 * left-subtree || left child digest || right child digest || right-subtree; returns the root.
 * leaf_digests are the hash_or_noop values of the leaves (hashed up-front so the dominant cost,
 * ceil(leaf_len/8) permutations per leaf, parallelises over all host threads). */
static void fill_subtree(uint64_t *digests_buf, size_t n_digests, const uint64_t *leaf_digests,
                         size_t n_leaves, uint64_t root[4]) {
  if (n_digests == 0) {
    memcpy(root, leaf_digests, 32);
    return;
  }
  size_t half = n_digests / 2;
  uint64_t *left_buf = digests_buf; /* half-1 digests */
  uint64_t *left_mem = digests_buf + 4 * (half - 1);
  uint64_t *right_mem = digests_buf + 4 * half;
  uint64_t *right_buf = digests_buf + 4 * (half + 1);
  uint64_t l[4], r[4];
  fill_subtree(left_buf, half - 1, leaf_digests, n_leaves / 2, l);
  fill_subtree(right_buf, half - 1, leaf_digests + 4 * (n_leaves / 2), n_leaves / 2, r);
  memcpy(left_mem, l, 32);
  memcpy(right_mem, r, 32);
  poseidon_two_to_one(l, r, root);
}

void merkle_tree_new(const uint64_t *leaves, size_t n_leaves, size_t leaf_len, unsigned cap_height,
                     uint64_t *digests_out, uint64_t *cap_out) {
  size_t n_cap = (size_t)1 << cap_height;
  size_t n_digests = 2 * (n_leaves - n_cap);
  size_t sub_digests = n_digests >> cap_height, sub_leaves = n_leaves >> cap_height;
  uint64_t *leaf_digests = (uint64_t *)malloc(n_leaves * 32);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n_leaves; i++)
    poseidon_hash_or_noop(leaves + i * leaf_len, leaf_len, leaf_digests + 4 * i);
#pragma omp parallel for schedule(dynamic)
  for (size_t i = 0; i < n_cap; i++)
    fill_subtree(digests_out + 4 * i * sub_digests, sub_digests, leaf_digests + 4 * i * sub_leaves,
                 sub_leaves, cap_out + 4 * i);
  free(leaf_digests);
}

void merkle_prove(const uint64_t *digests, size_t n_leaves, unsigned cap_height, size_t leaf_index,
                  uint64_t *siblings_out) {
  unsigned log_n = 0;
  while (((size_t)1 << log_n) < n_leaves) log_n++;
  unsigned num_layers = log_n - cap_height;
  size_t tree_len = (2 * (n_leaves - ((size_t)1 << cap_height))) >> cap_height;
  const uint64_t *digest_tree = digests + 4 * tree_len * (leaf_index >> num_layers);
  size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
  for (unsigned i = 0; i < num_layers; i++) {
    size_t parity = pair_index & 1;
    pair_index >>= 1;
    size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
    size_t sibling_index = 2 * siblings_index + (1 - parity);
    memcpy(siblings_out + 4 * i, digest_tree + 4 * sibling_index, 32);
  }
}

int merkle_verify(const uint64_t *leaf, size_t leaf_len, size_t leaf_index, const uint64_t *siblings,
                  unsigned n_siblings, const uint64_t *cap) {
  uint64_t cur[4], nxt[4];
  poseidon_hash_or_noop(leaf, leaf_len, cur);
  size_t index = leaf_index;
  for (unsigned i = 0; i < n_siblings; i++) {
    if (index & 1)
      poseidon_two_to_one(siblings + 4 * i, cur, nxt);
    else
      poseidon_two_to_one(cur, siblings + 4 * i, nxt);
    memcpy(cur, nxt, 32);
    index >>= 1;
  }
  return memcmp(cur, cap + 4 * index, 32) == 0;
}

/* Test helper for the stored proofs of qbench_data/example.bin: a FRI query index is derived from the
 * transcript (not available offline), so recover it from the path itself: try every left/right
 * pattern and return the unique leaf index whose root equals cap[index >> n_siblings], or -1 if
 * none / -2 if ambiguous. */
long merkle_find_index(const uint64_t *leaf, size_t leaf_len, const uint64_t *siblings, unsigned n_siblings,
                       const uint64_t *cap, size_t n_cap) {
  size_t n_cand = (size_t)1 << n_siblings;
  uint64_t *cur = (uint64_t *)malloc(n_cand * 32), *nxt = (uint64_t *)malloc(n_cand * 32);
  poseidon_hash_or_noop(leaf, leaf_len, cur);
  /* after level i, cur[b] is the node value assuming the low i+1 index bits are b (bit i = MSB of b) */
  size_t cnt = 1;
  for (unsigned i = 0; i < n_siblings; i++) {
    for (size_t b = 0; b < cnt; b++) {
      poseidon_two_to_one(cur + 4 * b, siblings + 4 * i, nxt + 4 * b);         /* bit i = 0 */
      poseidon_two_to_one(siblings + 4 * i, cur + 4 * b, nxt + 4 * (b + cnt)); /* bit i = 1 */
    }
    cnt *= 2;
    uint64_t *t = cur;
    cur = nxt;
    nxt = t;
  }
  long found = -1;
  for (size_t b = 0; b < cnt; b++)
    for (size_t c = 0; c < n_cap; c++)
      if (memcmp(cur + 4 * b, cap + 4 * c, 32) == 0) {
        long idx = (long)((c << n_siblings) | b);
        found = (found == -1) ? idx : -2;
      }
  free(cur);
  free(nxt);
  return found;
}
