// Host-side thread setting of the drop-in (reference: settings::get_n_threads / set_n_threads / reset_n_threads,
// include/piranha/settings.hpp:102-128).  In the reference it sizes the multiplier's thread pool; here the product
// itself runs on the GPU and the setting only sizes the host passes around it (the fill of the result container).
#ifndef PB200_SETTINGS_HPP
#define PB200_SETTINGS_HPP

#include <atomic>
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

public:
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
