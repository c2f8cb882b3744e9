/* TEST INFRASTRUCTURE ONLY - see oracle.h.  Reference-sequence fetch from the 2-bit forward strand. */
#include "oracle.h"

static int pac_base(const uint8_t *pac, int64_t p) { return (pac[p >> 2] >> (2 * (3 - (p & 3)))) & 3; }

/* bns_get_seq, reference src/bntseq.c:398-419: [beg,end) in the forward+reverse-complement coordinate [0, 2*l_pac);
 * a window bridging the strand boundary yields nothing; reverse-strand windows come out reverse-complemented. */
int64_t orc_get_seq(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, uint8_t *out)
{
	int64_t k, n = 0;
	if (end < beg) { int64_t t = beg; beg = end; end = t; }
	if (end > 2 * l_pac) end = 2 * l_pac;
	if (beg < 0) beg = 0;
	if (beg >= l_pac || end <= l_pac) {
		if (beg >= l_pac) {
			int64_t fb = 2 * l_pac - 1 - beg, fe = 2 * l_pac - 1 - end;   /* walk the forward strand downwards */
			for (k = fb; k > fe; --k) out[n++] = 3 - pac_base(pac, k);
		} else for (k = beg; k < end; ++k) out[n++] = pac_base(pac, k);
	}
	return n;
}
