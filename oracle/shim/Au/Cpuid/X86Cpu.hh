/* TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Stand-in for the one un-vendored third-party header the reference's hot path
 * includes: aoclutils' "Au/Cpuid/X86Cpu.hh" (found via find_library(aoclutils),
 * cmake/Dependencies.cmake:93-98 of the reference; no version is pinned in-tree).
 * The reference uses it in exactly one place, context::context()
 * (library/src/include/aoclsparse_context.hpp:142-250), to ask which ISA flags
 * the host CPU has and which Zen generation it is. No arithmetic lives there.
 * This shim answers the flag queries with the compiler builtin and reports an
 * unknown micro-architecture, which makes the reference fall back to its
 * "map by AVX flags" branch (aoclsparse_context.hpp:231-249).
 */
#pragma once
#include <climits>
#include <cmath>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

namespace Au
{
    enum class EUarch
    {
        Unknown = 0,
        Zen,
        ZenPlus,
        Zen2,
        Zen3,
        Zen4,
        Zen5
    };

    enum class ECpuidFlag
    {
        avx2,
        avx512f,
        avx512dq,
        avx512vl,
        avx512ifma,
        avx512cd,
        avx512bw,
        avx512vbmi,
        avx512_4vnniw,
        avx512_vpopcntdq
    };

    class X86Cpu
    {
    public:
        X86Cpu(int = 0)
        {
            __builtin_cpu_init();
        }
        EUarch getUarch() const
        {
            return EUarch::Unknown;
        }
        bool hasFlag(ECpuidFlag f) const
        {
            switch(f)
            {
            case ECpuidFlag::avx2:
                return __builtin_cpu_supports("avx2");
            case ECpuidFlag::avx512f:
                return __builtin_cpu_supports("avx512f");
            case ECpuidFlag::avx512dq:
                return __builtin_cpu_supports("avx512dq");
            case ECpuidFlag::avx512vl:
                return __builtin_cpu_supports("avx512vl");
            case ECpuidFlag::avx512ifma:
                return __builtin_cpu_supports("avx512ifma");
            case ECpuidFlag::avx512cd:
                return __builtin_cpu_supports("avx512cd");
            case ECpuidFlag::avx512bw:
                return __builtin_cpu_supports("avx512bw");
            case ECpuidFlag::avx512vbmi:
                return __builtin_cpu_supports("avx512vbmi");
            case ECpuidFlag::avx512_4vnniw:
                return __builtin_cpu_supports("avx5124vnniw");
            case ECpuidFlag::avx512_vpopcntdq:
                return __builtin_cpu_supports("avx512vpopcntdq");
            }
            return false;
        }
    };
}
