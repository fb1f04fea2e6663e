// rg_rendezvous.h — agreement among the worker threads of one rg_group call (rg_group.cu), host-only C++11.
//
// Every worker drives one GPU and enters the same sequence of NCCL collectives.  A worker that fails BEFORE a collective
// (an allocation, an upload) must not leave its peers blocked inside it, so the workers meet first: agree(ok) returns the
// logical AND of every participant's `ok` to all of them, and they proceed to the collective only when it is true.  The
// object is reusable (one generation per meeting); all n participants must call agree() the same number of times.
#pragma once
#include <condition_variable>
#include <mutex>

namespace rg {

class Rendezvous {
  public:
    explicit Rendezvous(int n) : n_(n) {}
    bool agree(bool ok) {
        std::unique_lock<std::mutex> lk(m_);
        const unsigned gen = gen_;
        all_ok_ = all_ok_ && ok;
        if (++arrived_ == n_) {
            result_ = all_ok_;
            arrived_ = 0;
            all_ok_ = true;
            ++gen_;
            cv_.notify_all();
            return result_;
        }
        cv_.wait(lk, [&] { return gen_ != gen; });
        return result_;  // not overwritten before every waiter of this generation has left: the next meeting needs all n again
    }

  private:
    const int n_;
    std::mutex m_;
    std::condition_variable cv_;
    int arrived_ = 0;
    unsigned gen_ = 0;
    bool all_ok_ = true, result_ = true;
};

}  // namespace rg
