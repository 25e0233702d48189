// pwgs_host_pickle.cpp — host code, no device: byte assembly for the tree pickles evolve.py writes into trees.zip
// (evolve.py:233 cPickle.dumps(tssb, protocol=2); format read back by util2.py:264-343).  host/fastpickle.py keeps one constant
// byte fragment per datum; what changes from tree to tree is only which node each datum refers to.  Joining the 50k fragments
// with their node references was the largest Python-level cost of a tree sample after the sweep; it is a gather of byte
// strings, done here in one pass.
#include "../../include/pwgs_b200.h"
#include <cstring>

extern "C" int64_t pwgs_pickle_join(int32_t n, const uint8_t *frags, const int64_t *frag_off, const int32_t *ref_of, int32_t n_refs,
                                    const uint8_t *refs, const uint8_t *tail, int32_t tail_len, uint8_t *out, int64_t cap)
{
    if (n < 0 || !frag_off || (n > 0 && (!frags || !ref_of || !out)) || tail_len < 0 || (tail_len > 0 && !tail)) return PWGS_EINVAL;
    int64_t w = 0;
    for (int32_t i = 0; i < n; i++) {
        const int64_t lo = frag_off[i], len = frag_off[i + 1] - lo;
        if (len < 0) return PWGS_EINVAL;
        const int32_t r = ref_of[i];
        if (r >= n_refs) return PWGS_EINVAL;
        const int64_t need = len + (r >= 0 ? (int64_t)refs[8 * (int64_t)r] + tail_len : 0);
        if (w + need > cap) return PWGS_ELIMIT;
        std::memcpy(out + w, frags + lo, (size_t)len);
        w += len;
        if (r >= 0) {                                         // r < 0: the fragment stands alone (a back-reference to a datum defined earlier)
            const uint8_t *ref = refs + 8 * (int64_t)r;       // byte 0: length (<= 7), then the opcode bytes of the reference
            if (ref[0] > 7) return PWGS_EINVAL;
            std::memcpy(out + w, ref + 1, ref[0]);
            w += ref[0];
            std::memcpy(out + w, tail, (size_t)tail_len);
            w += tail_len;
        }
    }
    return w;
}
