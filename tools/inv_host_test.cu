// Host-side check of csrc/inv_bin.cuh (no GPU needed): reads "a p" hex pairs on stdin, prints a^-1 mod p.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "../zk-fhe_b200/csrc/inv_bin.cuh"

static void parse(const char* hex, uint32_t* out) {
    char buf[65];
    size_t len = strlen(hex);
    memset(buf, '0', 64);
    buf[64] = 0;
    memcpy(buf + 64 - len, hex, len);
    for (int i = 0; i < 8; i++) {
        char w[9];
        memcpy(w, buf + 64 - 8 * (i + 1), 8);
        w[8] = 0;
        out[i] = (uint32_t)strtoul(w, nullptr, 16);
    }
}

int main() {
    char a_hex[80], p_hex[80];
    while (scanf("%79s %79s", a_hex, p_hex) == 2) {
        uint32_t a[8], p[8], r[8];
        parse(a_hex, a);
        parse(p_hex, p);
        zkfhe::u256_inv_odd(r, a, p);
        for (int i = 7; i >= 0; i--) printf("%08x", r[i]);
        printf("\n");
    }
    return 0;
}
