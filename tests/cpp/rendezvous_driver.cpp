// Driver of tests/test_host_types.py::test_rendezvous_*: n threads meet `rounds` times; in round r thread (r % n) reports a
// failure when r is a multiple of `fail_every`.  Every thread must see the same verdict in every round, and that verdict must
// be "ok" exactly in the rounds nobody failed.  Prints "ok <rounds> <failed rounds>" or "MISMATCH ...".
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "rg_rendezvous.h"

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 4, rounds = argc > 2 ? atoi(argv[2]) : 1000, fail_every = argc > 3 ? atoi(argv[3]) : 7;
    rg::Rendezvous rv(n);
    std::vector<std::vector<char>> seen(n, std::vector<char>(rounds));
    std::vector<std::thread> th;
    for (int t = 0; t < n; ++t)
        th.emplace_back([&, t] {
            for (int r = 0; r < rounds; ++r) {
                const bool ok = !(r % fail_every == 0 && r % n == t);
                if ((r + t) % 3 == 0) std::this_thread::yield();
                seen[t][r] = rv.agree(ok) ? 1 : 0;
            }
        });
    for (auto &x : th) x.join();
    int failed = 0;
    for (int r = 0; r < rounds; ++r) {
        const char want = (r % fail_every == 0) ? 0 : 1;
        failed += want == 0;
        for (int t = 0; t < n; ++t)
            if (seen[t][r] != want) {
                printf("MISMATCH round %d thread %d saw %d, expected %d\n", r, t, seen[t][r], want);
                return 1;
            }
    }
    printf("ok %d %d\n", rounds, failed);
    return 0;
}
