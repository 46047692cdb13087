/* zb_entropy.cuh — device-side entropy table builders (Huffman + FSE).
 *
 * These are the small, inherently serial steps of the entropy stage (<=256 / <=53 symbols).  They
 * run on one thread of the block's CTA, out of shared memory, never on the host.  Each function
 * names the reference code whose *result* it must reproduce bit-for-bit
 * (paths relative to /root/reference/lib); the bulk encoders that use the tables are the
 * parallel kernels in zb_literals.cu / zb_sequences.cu.
 */
#ifndef ZB_ENTROPY_CUH
#define ZB_ENTROPY_CUH
#include "zb_device.cuh"

#define ZBD_ERR 0xFFFFFFFFu           /* "could not build" -> caller falls back to raw/basic */

/* ------------------------------------------------------------------ serial LE bit writer */
struct ZbdBitW { u8* out; u32 pos; u64 acc; u32 nacc; };
__device__ __forceinline__ void zbd_bw_init(ZbdBitW* w, u8* out) { w->out = out; w->pos = 0; w->acc = 0; w->nacc = 0; }
__device__ __forceinline__ void zbd_bw_add(ZbdBitW* w, u32 value, u32 nbBits)
{
    if (!nbBits) return;
    w->acc |= (u64)(value & ((1u << nbBits) - 1u)) << w->nacc;       /* nbBits <= 16 here */
    w->nacc += nbBits;
    while (w->nacc >= 8) { w->out[w->pos++] = (u8)w->acc; w->acc >>= 8; w->nacc -= 8; }
}
__device__ __forceinline__ u32 zbd_bw_close(ZbdBitW* w)              /* common/bitstream.h:235-241 */
{
    zbd_bw_add(w, 1, 1);
    if (w->nacc) w->out[w->pos++] = (u8)w->acc;
    return w->pos;
}

/* ------------------------------------------------------------------ FSE */
/* ZbdFseCTable: zb_common.h */

/* compress/fse_compress.c:347-374 */
__device__ __forceinline__ u32 zbd_fse_minTableLog(u32 srcSize, u32 maxSymbolValue)
{
    u32 const a = zb_hb32(srcSize) + 1, b = zb_hb32(maxSymbolValue) + 2;
    return a < b ? a : b;
}
__device__ __forceinline__ u32 zbd_fse_optimalTableLog(u32 maxTableLog, u32 srcSize, u32 maxSymbolValue, u32 minus)
{
    u32 const maxBitsSrc = zb_hb32(srcSize - 1) - minus;
    u32 tableLog = maxTableLog;
    u32 const minBits = zbd_fse_minTableLog(srcSize, maxSymbolValue);
    if (maxBitsSrc < tableLog) tableLog = maxBitsSrc;
    if (minBits > tableLog) tableLog = minBits;
    if (tableLog < 5) tableLog = 5;
    if (tableLog > 12) tableLog = 12;
    return tableLog;
}

/* compress/fse_compress.c:379-463 */
__device__ inline u32 zbd_fse_normalizeM2(short* norm, u32 tableLog, const u32* count, u32 total, u32 maxSymbolValue, short lowProbCount)
{
    short const NOT_YET = -2;
    u32 s, distributed = 0, toDistribute;
    u32 const lowThreshold = total >> tableLog;
    u32 lowOne = (u32)(((u64)total * 3) >> (tableLog + 1));
    for (s = 0; s <= maxSymbolValue; s++) {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = lowProbCount; distributed++; total -= count[s]; continue; }
        if (count[s] <= lowOne) { norm[s] = 1; distributed++; total -= count[s]; continue; }
        norm[s] = NOT_YET;
    }
    toDistribute = (1u << tableLog) - distributed;
    if (toDistribute == 0) return 0;
    if ((total / toDistribute) > lowOne) {
        lowOne = (u32)(((u64)total * 3) / (toDistribute * 2));
        for (s = 0; s <= maxSymbolValue; s++)
            if (norm[s] == NOT_YET && count[s] <= lowOne) { norm[s] = 1; distributed++; total -= count[s]; }
        toDistribute = (1u << tableLog) - distributed;
    }
    if (distributed == maxSymbolValue + 1) {
        u32 maxV = 0, maxC = 0;
        for (s = 0; s <= maxSymbolValue; s++) if (count[s] > maxC) { maxV = s; maxC = count[s]; }
        norm[maxV] += (short)toDistribute;
        return 0;
    }
    if (total == 0) {
        for (s = 0; toDistribute > 0; s = (s + 1) % (maxSymbolValue + 1))
            if (norm[s] > 0) { toDistribute--; norm[s]++; }
        return 0;
    }
    {   u64 const vStepLog = 62 - tableLog;
        u64 const mid = (1ull << (vStepLog - 1)) - 1;
        u64 const rStep = ((((u64)1 << vStepLog) * toDistribute) + mid) / total;
        u64 tmpTotal = mid;
        for (s = 0; s <= maxSymbolValue; s++) {
            if (norm[s] == NOT_YET) {
                u64 const end = tmpTotal + ((u64)count[s] * rStep);
                u32 const sStart = (u32)(tmpTotal >> vStepLog);
                u32 const sEnd = (u32)(end >> vStepLog);
                u32 const weight = sEnd - sStart;
                if (weight < 1) return ZBD_ERR;
                norm[s] = (short)weight;
                tmpTotal = end;
            }
        }
    }
    return 0;
}

/* compress/fse_compress.c:465-525.  Returns tableLog, 0 for the rle special case, ZBD_ERR on failure. */
__device__ inline u32 zbd_fse_normalize(short* norm, u32 tableLog, const u32* count, u32 total, u32 maxSymbolValue, u32 useLowProbCount)
{
    const u32 rtb[8] = { 0, 473195, 504333, 520860, 550000, 700000, 750000, 830000 };
    if (tableLog < 5 || tableLog > 12) return ZBD_ERR;
    if (tableLog < zbd_fse_minTableLog(total, maxSymbolValue)) return ZBD_ERR;
    short const lowProbCount = useLowProbCount ? (short)-1 : (short)1;
    u64 const scale = 62 - tableLog;
    u64 const step = ((u64)1 << 62) / total;
    u64 const vStep = 1ull << (scale - 20);
    int stillToDistribute = 1 << tableLog;
    u32 s, largest = 0;
    short largestP = 0;
    u32 const lowThreshold = total >> tableLog;
    for (s = 0; s <= maxSymbolValue; s++) {
        if (count[s] == total) return 0;
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = lowProbCount; stillToDistribute--; }
        else {
            short proba = (short)(((u64)count[s] * step) >> scale);
            if (proba < 8) {
                u64 const restToBeat = vStep * rtb[proba];
                proba += (((u64)count[s] * step) - ((u64)proba << scale)) > restToBeat;
            }
            if (proba > largestP) { largestP = proba; largest = s; }
            norm[s] = proba;
            stillToDistribute -= proba;
        }
    }
    if (-stillToDistribute >= (norm[largest] >> 1)) {
        if (zbd_fse_normalizeM2(norm, tableLog, count, total, maxSymbolValue, lowProbCount) == ZBD_ERR) return ZBD_ERR;
    } else norm[largest] += (short)stillToDistribute;
    return tableLog;
}

/* compress/fse_compress.c:234-327.  Output capacity is the caller's business (<= 133 bytes needed). */
__device__ inline u32 zbd_fse_writeNCount(u8* dst, const short* norm, u32 maxSymbolValue, u32 tableLog)
{
    u8* out = dst;
    int nbBits;
    int const tableSize = 1 << tableLog;
    int remaining, threshold;
    u32 bitStream = 0;
    int bitCount = 0;
    u32 symbol = 0;
    u32 const alphabetSize = maxSymbolValue + 1;
    int previousIs0 = 0;

    bitStream += (tableLog - 5) << bitCount;
    bitCount += 4;
    remaining = tableSize + 1;
    threshold = tableSize;
    nbBits = (int)tableLog + 1;
    while (symbol < alphabetSize && remaining > 1) {
        if (previousIs0) {
            u32 start = symbol;
            while (symbol < alphabetSize && !norm[symbol]) symbol++;
            if (symbol == alphabetSize) break;
            while (symbol >= start + 24) {
                start += 24;
                bitStream += 0xFFFFu << bitCount;
                out[0] = (u8)bitStream; out[1] = (u8)(bitStream >> 8); out += 2; bitStream >>= 16;
            }
            while (symbol >= start + 3) { start += 3; bitStream += 3u << bitCount; bitCount += 2; }
            bitStream += (symbol - start) << bitCount;
            bitCount += 2;
            if (bitCount > 16) { out[0] = (u8)bitStream; out[1] = (u8)(bitStream >> 8); out += 2; bitStream >>= 16; bitCount -= 16; }
        }
        {   int count = norm[symbol++];
            int const max = (2 * threshold - 1) - remaining;
            remaining -= count < 0 ? -count : count;
            count++;
            if (count >= threshold) count += max;
            bitStream += (u32)count << bitCount;
            bitCount += nbBits;
            bitCount -= (count < max);
            previousIs0 = (count == 1);
            if (remaining < 1) return ZBD_ERR;
            while (remaining < threshold) { nbBits--; threshold >>= 1; }
        }
        if (bitCount > 16) { out[0] = (u8)bitStream; out[1] = (u8)(bitStream >> 8); out += 2; bitStream >>= 16; bitCount -= 16; }
    }
    if (remaining != 1) return ZBD_ERR;
    out[0] = (u8)bitStream;
    out[1] = (u8)(bitStream >> 8);
    out += (bitCount + 7) / 8;
    return (u32)(out - dst);
}

/* compress/fse_compress.c:68-214.  `scratch` >= 512 + 134 bytes. */
__device__ inline void zbd_fse_buildCTable(ZbdFseCTable* ct, const short* norm, u32 maxSymbolValue, u32 tableLog, u8* scratch)
{
    u32 const tableSize = 1u << tableLog;
    u32 const tableMask = tableSize - 1;
    u32 const step = (tableSize >> 1) + (tableSize >> 3) + 3;        /* common/fse.h:632 */
    u32 const maxSV1 = maxSymbolValue + 1;
    u8*  const tableSymbol = scratch;
    u16* const cumul = (u16*)(scratch + 512);
    u32 highThreshold = tableSize - 1;
    ct->tableLog = tableLog;
    ct->maxSymbolValue = maxSymbolValue;
    cumul[0] = 0;
    for (u32 u = 1; u <= maxSV1; u++) {
        if (norm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1; tableSymbol[highThreshold--] = (u8)(u - 1); }
        else cumul[u] = cumul[u - 1] + (u16)norm[u - 1];
    }
    cumul[maxSV1] = (u16)(tableSize + 1);
    {   u32 position = 0;
        for (u32 symbol = 0; symbol < maxSV1; symbol++) {
            int const freq = norm[symbol];
            for (int i = 0; i < freq; i++) {
                tableSymbol[position] = (u8)symbol;
                position = (position + step) & tableMask;
                while (position > highThreshold) position = (position + step) & tableMask;
            }
        }
    }
    for (u32 u = 0; u < tableSize; u++) { u8 const s = tableSymbol[u]; ct->nextState[cumul[s]++] = (u16)(tableSize + u); }
    {   u32 total = 0;
        for (u32 s = 0; s <= maxSymbolValue; s++) {
            int const n = norm[s];
            if (n == 0) { ct->deltaNbBits[s] = ((tableLog + 1) << 16) - (1u << tableLog); ct->deltaFindState[s] = 0; }
            else if (n == -1 || n == 1) { ct->deltaNbBits[s] = (tableLog << 16) - (1u << tableLog); ct->deltaFindState[s] = (int)(total - 1); total++; }
            else {
                u32 const maxBitsOut = tableLog - zb_hb32((u32)n - 1);
                u32 const minStatePlus = (u32)n << maxBitsOut;
                ct->deltaNbBits[s] = (maxBitsOut << 16) - minStatePlus;
                ct->deltaFindState[s] = (int)(total - (u32)n);
                total += (u32)n;
            }
        }
    }
}
/* compress/fse_compress.c:528-549 */
__device__ inline void zbd_fse_buildCTable_rle(ZbdFseCTable* ct, u32 symbol)
{
    ct->tableLog = 0; ct->maxSymbolValue = symbol;
    ct->nextState[0] = 0; ct->nextState[1] = 0;
    ct->deltaNbBits[symbol] = 0; ct->deltaFindState[symbol] = 0;
}

/* common/fse.h:452-476 */
__device__ __forceinline__ u32 zbd_fse_initState2(const ZbdFseCTable* ct, u32 symbol)
{
    u32 const dnb = ct->deltaNbBits[symbol];
    u32 const nbBitsOut = (dnb + (1u << 15)) >> 16;
    u32 const value = (nbBitsOut << 16) - dnb;
    return ct->nextState[(int)(value >> nbBitsOut) + ct->deltaFindState[symbol]];
}
/* returns next state; *bits / *nb receive the emitted field */
__device__ __forceinline__ u32 zbd_fse_step(const ZbdFseCTable* ct, u32 state, u32 symbol, u32* bits, u32* nb)
{
    u32 const nbBitsOut = (state + ct->deltaNbBits[symbol]) >> 16;
    *nb = nbBitsOut;
    *bits = state & ((1u << nbBitsOut) - 1u);
    return ct->nextState[(int)(state >> nbBitsOut) + ct->deltaFindState[symbol]];
}

/* ------------------------------------------------------------------ Huffman table build */
struct ZbdHNode { u32 count; u16 parent; u8 byte; u8 nbBits; };

#define ZBD_RANK_TABLE 192
#define ZBD_RANK_LOG_BEGIN 158
#define ZBD_RANK_CUTOFF 166
__device__ __forceinline__ u32 zbd_huf_bucket(u32 c) { return c < ZBD_RANK_CUTOFF ? c : zb_hb32(c) + ZBD_RANK_LOG_BEGIN; }

/* compress/huf_compress.c:564-615 — the reference's (unstable) sort decides ties, so the exact
 * partition scheme is kept; recursion is replaced by an explicit stack of pending ranges. */
__device__ inline void zbd_huf_insertionSort(ZbdHNode* arr, int low, int high)
{
    int const size = high - low + 1;
    arr += low;
    for (int i = 1; i < size; i++) {
        ZbdHNode const key = arr[i];
        int j = i - 1;
        while (j >= 0 && arr[j].count < key.count) { arr[j + 1] = arr[j]; j--; }
        arr[j + 1] = key;
    }
}
__device__ inline void zbd_huf_quickSort(ZbdHNode* arr, int low0, int high0, short* stack /* >= 2*260 */)
{
    int sp = 0;
    stack[sp++] = (short)low0; stack[sp++] = (short)high0;
    while (sp > 0) {
        int high = stack[--sp];
        int low = stack[--sp];
        if (high - low < 8) { zbd_huf_insertionSort(arr, low, high); continue; }
        while (low < high) {
            u32 const pivot = arr[high].count;
            int i = low - 1;
            for (int j = low; j < high; j++) if (arr[j].count > pivot) { i++; ZbdHNode t = arr[i]; arr[i] = arr[j]; arr[j] = t; }
            { ZbdHNode t = arr[i + 1]; arr[i + 1] = arr[high]; arr[high] = t; }
            int const idx = i + 1;
            if (idx - low < high - idx) { stack[sp++] = (short)low; stack[sp++] = (short)(idx - 1); low = idx + 1; }
            else { stack[sp++] = (short)(idx + 1); stack[sp++] = (short)high; high = idx - 1; }
        }
    }
}

/* compress/huf_compress.c:376-497 */
__device__ inline u32 zbd_huf_setMaxHeight(ZbdHNode* node, u32 lastNonNull, u32 targetNbBits)
{
    u32 const largestBits = node[lastNonNull].nbBits;
    if (largestBits <= targetNbBits) return largestBits;
    int totalCost = 0;
    u32 const baseCost = 1u << (largestBits - targetNbBits);
    int n = (int)lastNonNull;
    while (node[n].nbBits > targetNbBits) {
        totalCost += (int)(baseCost - (1u << (largestBits - node[n].nbBits)));
        node[n].nbBits = (u8)targetNbBits;
        n--;
    }
    while (node[n].nbBits == targetNbBits) --n;
    totalCost >>= (largestBits - targetNbBits);
    u32 const noSymbol = 0xF0F0F0F0u;
    u32 rankLast[14];
    for (int i = 0; i < 14; i++) rankLast[i] = noSymbol;
    {   u32 currentNbBits = targetNbBits;
        for (int pos = n; pos >= 0; pos--) {
            if (node[pos].nbBits >= currentNbBits) continue;
            currentNbBits = node[pos].nbBits;
            rankLast[targetNbBits - currentNbBits] = (u32)pos;
        }
    }
    while (totalCost > 0) {
        u32 nBitsToDecrease = zb_hb32((u32)totalCost) + 1;
        for (; nBitsToDecrease > 1; nBitsToDecrease--) {
            u32 const highPos = rankLast[nBitsToDecrease];
            u32 const lowPos = rankLast[nBitsToDecrease - 1];
            if (highPos == noSymbol) continue;
            if (lowPos == noSymbol) break;
            if (node[highPos].count <= 2 * node[lowPos].count) break;
        }
        while (nBitsToDecrease <= 12 && rankLast[nBitsToDecrease] == noSymbol) nBitsToDecrease++;
        totalCost -= 1 << (nBitsToDecrease - 1);
        node[rankLast[nBitsToDecrease]].nbBits++;
        if (rankLast[nBitsToDecrease - 1] == noSymbol) rankLast[nBitsToDecrease - 1] = rankLast[nBitsToDecrease];
        if (rankLast[nBitsToDecrease] == 0) rankLast[nBitsToDecrease] = noSymbol;
        else {
            rankLast[nBitsToDecrease]--;
            if (node[rankLast[nBitsToDecrease]].nbBits != targetNbBits - nBitsToDecrease) rankLast[nBitsToDecrease] = noSymbol;
        }
    }
    while (totalCost < 0) {
        if (rankLast[1] == noSymbol) {
            while (node[n].nbBits == targetNbBits) n--;
            node[n + 1].nbBits--;
            rankLast[1] = (u32)(n + 1);
            totalCost++;
            continue;
        }
        node[rankLast[1] + 1].nbBits--;
        rankLast[1]++;
        totalCost++;
    }
    return targetNbBits;
}

/* Workspace the Huffman builder needs (shared memory, one per CTA). */
struct ZbdHufWksp {
    ZbdHNode table[2 * 256 + 1];     /* [0] is the sentinel of huf_compress.c:695 */
    u16 rankBase[ZBD_RANK_TABLE];
    u16 rankCurr[ZBD_RANK_TABLE];
    short stack[2 * 260];
    u8  weights[256];
    u8  fseScratch[512 + 136];
    ZbdFseCTable wct;                /* FSE table for the weights */
};

/* compress/huf_compress.c:756-791 (sort :620-668, tree :681-723, height :376, codes :730-753).
 * Fills enc[s] = code | nbBits << 16 for every symbol; returns the table's max code length. */
__device__ inline u32 zbd_huf_build(ZbdHufWksp* w, const u32* count, u32 maxSymbolValue, u32 maxNbBits, u32* enc)
{
    ZbdHNode* const node = w->table + 1;
    int const STARTNODE = 256;
    u32 const maxSV1 = maxSymbolValue + 1;
    for (u32 i = 0; i < 2 * 256 + 1; i++) { ZbdHNode z; z.count = 0; z.parent = 0; z.byte = 0; z.nbBits = 0; w->table[i] = z; }
    /* ---- sort ---- */
    for (u32 i = 0; i < ZBD_RANK_TABLE; i++) { w->rankBase[i] = 0; w->rankCurr[i] = 0; }
    for (u32 n = 0; n < maxSV1; n++) w->rankBase[zbd_huf_bucket(count[n])]++;
    for (u32 n = ZBD_RANK_TABLE - 1; n > 0; n--) { w->rankBase[n - 1] += w->rankBase[n]; w->rankCurr[n - 1] = w->rankBase[n - 1]; }
    for (u32 n = 0; n < maxSV1; n++) {
        u32 const c = count[n];
        u32 const r = zbd_huf_bucket(c) + 1;
        u32 const pos = w->rankCurr[r]++;
        node[pos].count = c; node[pos].byte = (u8)n;
    }
    for (u32 n = ZBD_RANK_CUTOFF; n < ZBD_RANK_TABLE - 1; n++) {
        int const bucketSize = (int)w->rankCurr[n] - (int)w->rankBase[n];
        if (bucketSize > 1) zbd_huf_quickSort(node + w->rankBase[n], 0, bucketSize - 1, w->stack);
    }
    /* ---- tree ---- */
    int nonNullRank = (int)maxSymbolValue;
    while (node[nonNullRank].count == 0) nonNullRank--;
    int lowS = nonNullRank, nodeNb = STARTNODE, lowN = STARTNODE;
    int const nodeRoot = nodeNb + lowS - 1;
    node[nodeNb].count = node[lowS].count + node[lowS - 1].count;
    node[lowS].parent = node[lowS - 1].parent = (u16)nodeNb;
    nodeNb++; lowS -= 2;
    for (int n = nodeNb; n <= nodeRoot; n++) node[n].count = 1u << 30;
    node[-1].count = 1u << 31;
    while (nodeNb <= nodeRoot) {
        int const n1 = (node[lowS].count < node[lowN].count) ? lowS-- : lowN++;
        int const n2 = (node[lowS].count < node[lowN].count) ? lowS-- : lowN++;
        node[nodeNb].count = node[n1].count + node[n2].count;
        node[n1].parent = node[n2].parent = (u16)nodeNb;
        nodeNb++;
    }
    node[nodeRoot].nbBits = 0;
    for (int n = nodeRoot - 1; n >= STARTNODE; n--) node[n].nbBits = node[node[n].parent].nbBits + 1;
    for (int n = 0; n <= nonNullRank; n++) node[n].nbBits = node[node[n].parent].nbBits + 1;
    /* ---- height limit + canonical codes ---- */
    maxNbBits = zbd_huf_setMaxHeight(node, (u32)nonNullRank, maxNbBits);
    if (maxNbBits > 12) return ZBD_ERR;
    u16 nbPerRank[13], valPerRank[13];
    for (int i = 0; i < 13; i++) { nbPerRank[i] = 0; valPerRank[i] = 0; }
    for (int n = 0; n <= nonNullRank; n++) nbPerRank[node[n].nbBits]++;
    {   u16 min = 0;
        for (int n = (int)maxNbBits; n > 0; n--) { valPerRank[n] = min; min += nbPerRank[n]; min >>= 1; }
    }
    for (u32 n = 0; n < 256; n++) enc[n] = 0;
    for (u32 n = 0; n < maxSV1; n++) enc[node[n].byte] = (u32)node[n].nbBits << 16;
    for (u32 n = 0; n < maxSV1; n++) {
        u32 const nb = enc[n] >> 16;
        if (nb) enc[n] |= valPerRank[nb]++;
    }
    return maxNbBits;
}

/* compress/huf_compress.c:248-289 + :147-186 (weights through a 2-state FSE, fse_compress.c:551-608).
 * Writes the tree description to dst (<= 129 bytes); returns its size or ZBD_ERR. */
__device__ inline u32 zbd_huf_writeHeader(ZbdHufWksp* w, u8* dst, const u32* enc, u32 maxSymbolValue, u32 huffLog)
{
    u8* const wt = w->weights;
    for (u32 n = 0; n < maxSymbolValue; n++) { u32 const nb = enc[n] >> 16; wt[n] = nb ? (u8)(huffLog + 1 - nb) : 0; }
    /* ---- HUF_compressWeights ---- */
    u32 hSize = 0;
    u32 const wtSize = maxSymbolValue;
    if (wtSize > 1) {
        u32 count[13]; short norm[13];
        for (int i = 0; i < 13; i++) count[i] = 0;
        for (u32 i = 0; i < wtSize; i++) count[wt[i]]++;
        u32 maxSym = 12; while (!count[maxSym]) maxSym--;
        u32 maxCount = 0; for (u32 s = 0; s <= maxSym; s++) if (count[s] > maxCount) maxCount = count[s];
        if (maxCount == wtSize) hSize = 1;
        else if (maxCount == 1) hSize = 0;
        else {
            u32 const tableLog = zbd_fse_optimalTableLog(6, wtSize, maxSym, 2);
            if (zbd_fse_normalize(norm, tableLog, count, wtSize, maxSym, 0) == ZBD_ERR) return ZBD_ERR;
            u8* op = dst + 1;
            u32 const nc = zbd_fse_writeNCount(op, norm, maxSym, tableLog);
            if (nc == ZBD_ERR) return ZBD_ERR;
            op += nc;
            zbd_fse_buildCTable(&w->wct, norm, maxSym, tableLog, w->fseScratch);
            if (wtSize <= 2) hSize = 0;
            else {
                ZbdBitW bw; zbd_bw_init(&bw, op);
                const u8* ip = wt + wtSize;
                u32 s1, s2, bits, nb;
                if (wtSize & 1) {
                    s1 = zbd_fse_initState2(&w->wct, *--ip);
                    s2 = zbd_fse_initState2(&w->wct, *--ip);
                    s1 = zbd_fse_step(&w->wct, s1, *--ip, &bits, &nb); zbd_bw_add(&bw, bits, nb);
                } else {
                    s2 = zbd_fse_initState2(&w->wct, *--ip);
                    s1 = zbd_fse_initState2(&w->wct, *--ip);
                }
                while (ip > wt) {
                    s2 = zbd_fse_step(&w->wct, s2, *--ip, &bits, &nb); zbd_bw_add(&bw, bits, nb);
                    s1 = zbd_fse_step(&w->wct, s1, *--ip, &bits, &nb); zbd_bw_add(&bw, bits, nb);
                }
                zbd_bw_add(&bw, s2, tableLog);
                zbd_bw_add(&bw, s1, tableLog);
                u32 const c = zbd_bw_close(&bw);
                hSize = nc + c;
            }
        }
    }
    if (hSize > 1 && hSize < maxSymbolValue / 2) { dst[0] = (u8)hSize; return hSize + 1; }
    if (maxSymbolValue > 128) return ZBD_ERR;
    dst[0] = (u8)(128 + (maxSymbolValue - 1));
    wt[maxSymbolValue] = 0;
    for (u32 n = 0; n < maxSymbolValue; n += 2) dst[(n / 2) + 1] = (u8)((wt[n] << 4) + wt[n + 1]);
    return ((maxSymbolValue + 1) / 2) + 1;
}

#endif
