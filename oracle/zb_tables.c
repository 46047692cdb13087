/* zb_tables.c — oracle model of the PRODUCT's entropy-table builders (TEST INFRASTRUCTURE ONLY).
 *
 * The reference derives Huffman code lengths with HUF_buildCTable_wksp (huf_compress.c:756) and normalises FSE
 * counts with FSE_normalizeCount (fse_compress.c:465).  Only the RESULT TYPE is fixed by the format: any complete
 * prefix code with lengths <= tableLog, any normalised distribution that sums to 1 << tableLog with every present
 * symbol >= 1, decodes.  The product uses its own algorithms for both (zstd_b200/csrc/zb_entropy.cuh, written for a
 * warp / a CTA); this file is their plain-C statement, bit-exact with the kernels.  zb_entropy.c keeps the
 * restatement of the reference's algorithms (pinned byte-for-byte against the compiled reference); which of the
 * two the frame-level oracle uses is selected by zbo_entropy_model (1 = product, the default).
 */
#include <string.h>
#include "zb_oracle.h"

int zbo_entropy_model = 1;

/* ---- FSE normalisation: largest remainders ------------------------------------------------------------
 * base[s] = max(1, floor(count[s] * T / total)) for present symbols (T = 1 << tableLog); a symbol lifted to 1 has
 * remainder 0.  If the bases sum to less than T, the symbols with the largest remainders (ties: lower symbol
 * first) get one more each; if they sum to more (many lifted symbols), the largest base (ties: lower symbol
 * first) gives one back, as often as needed. */
size_t zbo_fse_normalize_lr(int16_t* norm, u32 tableLog, const u32* count, size_t total, u32 maxSymbolValue)
{
    u32 const T = 1u << tableLog;
    u32 base[256], rem[256];
    u32 s, sum = 0;
    for (s = 0; s <= maxSymbolValue; s++) {
        base[s] = 0; rem[s] = 0;
        if (count[s]) {
            u64 const x = (u64)count[s] * T;
            u32 const q = (u32)(x / total);
            if (q == 0) base[s] = 1; else { base[s] = q; rem[s] = (u32)(x % total); }
        }
        sum += base[s];
    }
    if (sum < T) {
        u32 need = T - sum;
        u8 taken[256];
        memset(taken, 0, sizeof(taken));
        while (need--) {
            u32 best = 0; int found = 0;
            for (s = 0; s <= maxSymbolValue; s++)
                if (count[s] && !taken[s] && (!found || rem[s] > rem[best])) { best = s; found = 1; }
            if (!found) return ZBO_ERR(ZBO_error_GENERIC);
            taken[best] = 1; base[best]++;
        }
    } else {
        u32 over = sum - T;
        while (over--) {
            u32 best = 0;
            for (s = 1; s <= maxSymbolValue; s++) if (base[s] > base[best]) best = s;
            if (base[best] < 2) return ZBO_ERR(ZBO_error_GENERIC);
            base[best]--;
        }
    }
    for (s = 0; s <= maxSymbolValue; s++) norm[s] = (int16_t)base[s];
    return tableLog;
}

/* ---- Huffman code lengths ------------------------------------------------------------------------------
 * 1. present symbols ranked by (count descending, symbol ascending);
 * 2. optimal lengths by the in-place algorithm of Moffat & Katajainen ("In-place calculation of minimum-redundancy
 *    codes", WADS 1995) over the ascending weights;
 * 3. if the deepest leaf is deeper than `target`: every deeper leaf is lifted to `target`, and for every unit of
 *    Kraft excess that creates, the deepest leaf above the bottom level is pushed one level down together with one
 *    leaf taken from the bottom level (the length-limiting step of deflate encoders);
 * 4. the histogram of lengths is dealt back out by rank: the most frequent symbols get the shortest codes.
 * Returns the largest length in use. */
size_t zbo_huf_lengths_mk(u8* nbBits, const u32* count, u32 maxSymbolValue, u32 target)
{
    u32 rankSym[256], A[256], nl[64];
    u32 nz = 0, s, i, maxLen;
    memset(nbBits, 0, 256);
    for (s = 0; s <= maxSymbolValue; s++) if (count[s]) {
        u32 r = 0, j;
        for (j = 0; j <= maxSymbolValue; j++) if (count[j] > count[s] || (count[j] == count[s] && j < s)) r += count[j] != 0;
        rankSym[r] = s; nz++;
    }
    if (nz == 0) return 0;
    if (nz == 1) { nbBits[rankSym[0]] = 1; return 1; }
    for (i = 0; i < nz; i++) A[i] = count[rankSym[nz - 1 - i]];          /* ascending weights */
    {   /* phase 1: A[k] becomes the parent index of internal node k, built in place */
        u32 root = 0, leaf = 2, next;
        A[0] += A[1];
        for (next = 1; next + 1 < nz; next++) {
            if (leaf >= nz || A[root] < A[leaf]) { A[next] = A[root]; A[root++] = next; } else A[next] = A[leaf++];
            if (leaf >= nz || (root < next && A[root] < A[leaf])) { A[next] += A[root]; A[root++] = next; } else A[next] += A[leaf++];
        }
        /* phase 2: parent indices -> depths of the internal nodes */
        A[nz - 2] = 0;
        for (next = nz - 2; next-- > 0; ) A[next] = A[A[next]] + 1;
        /* phase 3: depths of the leaves, deepest first */
        {   int avbl = 1, used = 0, depth = 0;
            int rt = (int)nz - 2, nx = (int)nz - 1;
            while (avbl > 0) {
                while (rt >= 0 && (int)A[rt] == depth) { used++; rt--; }
                while (avbl > used) { A[nx--] = (u32)depth; avbl--; }
                avbl = 2 * used; depth++; used = 0;
            }
        }
    }
    /* A[i] = length of the i-th lightest symbol: non-increasing in i */
    memset(nl, 0, sizeof(nl));
    maxLen = A[0];
    for (i = 0; i < nz; i++) nl[A[i] < 63 ? A[i] : 63]++;
    if (maxLen > target) {
        u32 l, K = 0, E;
        for (l = target + 1; l < 64; l++) { nl[target] += nl[l]; nl[l] = 0; }
        for (l = 1; l <= target; l++) K += nl[l] << (target - l);
        E = K - (1u << target);
        while (E--) {
            u32 b = target - 1;
            while (nl[b] == 0) b--;
            nl[b]--; nl[b + 1] += 2; nl[target]--;
        }
        maxLen = target;
    }
    {   u32 r = 0, l;
        for (l = 1; l <= maxLen; l++) for (i = 0; i < nl[l]; i++) nbBits[rankSym[r++]] = (u8)l;
    }
    while (nl[maxLen] == 0) maxLen--;
    return maxLen;
}
