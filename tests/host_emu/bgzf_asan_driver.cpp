// bgzf_asan_driver.cpp -- TEST INFRASTRUCTURE ONLY.  The emulated BGZF kernels (bgzf_emu.cpp) under AddressSanitizer + UBSan with exactly sized heap
// buffers: compute-sanitizer is not available on the GPU pool, so out-of-bounds indexing in the kernel bodies is looked for here.
//   g++ -O1 -g -fsanitize=address,undefined -o /tmp/bgzf_asan tests/host_emu/bgzf_asan_driver.cpp tests/host_emu/bgzf_emu.cpp -lz
//   ASAN_OPTIONS=detect_stack_use_after_return=0:detect_leaks=0 /tmp/bgzf_asan
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <zlib.h>
#include <stdint.h>
extern "C" long long bgzf_emu_deflate(const uint8_t *in, long long n_bytes, int level, uint8_t *out, int32_t *clen);
extern "C" int bgzf_emu_inflate_warp(const uint8_t *in, long long n_bytes, int n_members, const long long *member_off, uint8_t *out, long long *out_off, int *status);
extern "C" int bgzf_emu_inflate(const uint8_t *in, long long n_bytes, int n_members, const long long *member_off, uint8_t *out, long long *out_off, int *status);
int main()
{
	srand(5);
	for (int trial = 0; trial < 6; ++trial) {
		const size_t n = trial == 0 ? 1 : trial == 1 ? 65280 : trial == 2 ? 65281 : 30000 + rand() % 120000;
		std::vector<uint8_t> data(n);
		for (size_t i = 0; i < n; ++i) data[i] = trial == 3 ? (uint8_t)rand() : (uint8_t)("ACGTNacgt!!IIIIJJ"[rand() % 17] + (i % 300 < 20 ? i % 7 : 0));
		const size_t nblk = (n + 65279) / 65280;
		// exactly sized heap buffers so that the sanitizer sees any byte too many
		std::vector<uint8_t> out(nblk * 65536);
		std::vector<int32_t> clen(nblk);
		const long long tot = bgzf_emu_deflate(data.data(), (long long)n, 2, out.data(), clen.data());
		std::vector<long long> moff(nblk + 1, 0);
		for (size_t k = 0; k < nblk; ++k) moff[k + 1] = moff[k] + clen[k];
		if (moff[nblk] != tot) { printf("size mismatch\n"); return 1; }
		std::vector<uint8_t> packed(out.begin(), out.begin() + tot);
		std::vector<uint8_t> back(n);
		std::vector<long long> ooff(nblk + 1);
		std::vector<int> status(nblk);
		for (int form = 0; form < 2; ++form) {
			memset(back.data(), 0, n);
			const int bad = form ? bgzf_emu_inflate(packed.data(), tot, (int)nblk, moff.data(), back.data(), ooff.data(), status.data())
			                     : bgzf_emu_inflate_warp(packed.data(), tot, (int)nblk, moff.data(), back.data(), ooff.data(), status.data());
			if (bad || memcmp(back.data(), data.data(), n)) { printf("trial %d form %d: round trip failed (%d bad)\n", trial, form, bad); return 1; }
		}
		printf("trial %d: %zu bytes -> %lld, both inflate forms give them back\n", trial, n, tot);
	}
	return 0;
}
