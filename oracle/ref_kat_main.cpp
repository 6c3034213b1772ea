// TEST INFRASTRUCTURE. Mints known-answer vectors from the reference's OWN objects (linked from
// oracle/_ref/obj, compiled from /root/reference by oracle/Makefile): RollingHasherDual
// (source/hash_functions.cpp), next_prime3mod4 / modular inverse (source/functions_math.cpp),
// KMerFactoryCanonical2BC (source/kmer_factory.cpp).  Used only by tests/golden/make_golden.py.
//   ref_kat roll <q> <tbm 0|1> <ACGT string>   -> "Hf Hb fwd_is_canonical"  (window = whole string)
//   ref_kat prime <n>                          -> next_prime3mod4(n)
//   ref_kat inv <a> <m>                        -> modular inverse
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include "hash_functions.hpp"
#include "functions_math.hpp"
#include "functions_strings.hpp"
#include "kmer_factory.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    std::string cmd = argv[1];
    std::cout.setstate(std::ios_base::failbit);  // next_prime3mod4 chats on stdout
    if (cmd == "roll" && argc == 5) {
        uint64_t q = strtoull(argv[2], nullptr, 10);
        bool tbm = atoi(argv[3]) != 0;
        std::string s = argv[4];
        uint64_t k = s.size();
        uint64_t inv = mathfunctions::modular_multiplicative_inverse_coprimes(5, q);
        RollingHasherDual rh(q, k, inv, 5, 0, tbm);
        KMerFactoryCanonical2BC fac(k);
        for (char c : s) {
            uint64_t v = twobitstringfunctions::char2int(c);
            fac.push_new_integer(v);
            rh.update_rolling_hash(fac.get_forward_newest_character(), fac.get_forward_pushed_off_character());
        }
        printf("%llu %llu %d\n", (unsigned long long)rh.get_current_hash_forward_rqless(),
               (unsigned long long)rh.get_current_hash_backward_rqless(), fac.forward_kmer_is_canonical() ? 1 : 0);
    } else if (cmd == "prime") {
        printf("%llu\n", (unsigned long long)mathfunctions::next_prime3mod4(strtoull(argv[2], nullptr, 10)));
    } else if (cmd == "inv" && argc == 4) {
        printf("%llu\n", (unsigned long long)mathfunctions::modular_multiplicative_inverse_coprimes(
                             strtoll(argv[2], nullptr, 10), strtoll(argv[3], nullptr, 10)));
    } else return 2;
    return 0;
}
