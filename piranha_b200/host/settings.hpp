// Host-side thread setting of the drop-in (reference: settings::get_n_threads / set_n_threads / reset_n_threads,
// include/piranha/settings.hpp:102-128).  In the reference it sizes the multiplier's thread pool; here the product
// itself runs on the GPU and the setting only sizes the host passes around it (the fill of the result container).
#ifndef PB200_SETTINGS_HPP
#define PB200_SETTINGS_HPP

#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <thread>

namespace pb200
{

class settings
{
    static std::atomic<unsigned> &value()
    {
        static std::atomic<unsigned> v{hardware()};
        return v;
    }
    static unsigned hardware()
    {
        const unsigned c = std::thread::hardware_concurrency();
        return c ? c : 1u;
    }

    static std::atomic<unsigned> &gpus()
    {
        static std::atomic<unsigned> v{env_gpus()};
        return v;
    }
    static unsigned env_gpus()
    {
        if (const char *list = std::getenv("PB200_DEVICES")) { // "0,1,2,3": as many devices as the list names
            unsigned n = 1;
            for (const char *p = list; *p; ++p) n += (*p == ',');
            return n;
        }
        const char *e = std::getenv("PB200_N_GPUS");
        const int n = e ? std::atoi(e) : 1;
        return n > 0 ? static_cast<unsigned>(n) : 1u;
    }

public:
    // Number of GPUs one product is spread over (pk_ctx_create_multi): the multiplier's own parallelism, the role
    // set_n_threads plays for the reference's thread pool (settings.hpp:102-128).  Takes effect at the next product.
    static unsigned get_n_gpus() { return gpus().load(std::memory_order_relaxed); }
    static void set_n_gpus(unsigned n)
    {
        if (n == 0u || n > 16u) throw std::invalid_argument("the number of GPUs must be in [1, 16]");
        gpus().store(n, std::memory_order_relaxed);
    }
    static void reset_n_gpus() { gpus().store(env_gpus(), std::memory_order_relaxed); }
    static unsigned get_n_threads() { return value().load(std::memory_order_relaxed); }
    static void set_n_threads(unsigned n)
    {
        if (n == 0u) throw std::invalid_argument("the number of threads must be strictly positive");
        value().store(n, std::memory_order_relaxed);
    }
    static void reset_n_threads() { value().store(hardware(), std::memory_order_relaxed); }
};

} // namespace pb200

#endif
