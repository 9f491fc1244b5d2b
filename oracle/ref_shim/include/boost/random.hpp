// std:: equivalents: the VALUES differ from boost's, which is fine -- boost-derived values (fillers, CPU dropout
// masks) are parity-unpinned in the reference itself and are explicit inputs of every comparison.
#pragma once
#include <random>
#include "boost/random/mersenne_twister.hpp"
#include "boost/random/uniform_int.hpp"
namespace boost {
template <class T = double> class uniform_real : public std::uniform_real_distribution<T> { public: using std::uniform_real_distribution<T>::uniform_real_distribution; };
template <class T = double> class normal_distribution : public std::normal_distribution<T> { public: using std::normal_distribution<T>::normal_distribution; };
template <class T = double> class bernoulli_distribution { public: explicit bernoulli_distribution(T p = T(0.5)) : d_(double(p)) {} template <class G> int operator()(G& g) { return d_(g) ? 1 : 0; } typedef int result_type; private: std::bernoulli_distribution d_; };
template <class Engine, class Dist> class variate_generator {
 public:
  variate_generator(Engine e, Dist d) : e_(e), d_(d) {}
  typename Dist::result_type operator()() { return d_(*e_); }
 private:
  Engine e_; Dist d_;
};
}  // namespace boost
