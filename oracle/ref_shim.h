/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Force-included (-include) when compiling the UNMODIFIED reference sources
 * (/root/reference/src/cpp) with g++ into oracle/_ref/.  The reference is
 * written for MSVC, where <windows.h> supplies function-like macros
 *     #define max(a,b) (((a) > (b)) ? (a) : (b))
 * so calls such as max(1, some_double) compile.  With g++ only std::max<T>
 * exists and mixed-type calls fail to deduce.  These overloads give the mixed
 * calls the macro's semantics (usual arithmetic conversions); same-type calls
 * still bind to std::max / std::min.  No hot-path arithmetic is changed.
 */
#ifndef DTL_REF_SHIM_H
#define DTL_REF_SHIM_H
#ifdef __cplusplus
#include <algorithm>
#include <type_traits>

template <class A, class B,
          class = typename std::enable_if<!std::is_same<A, B>::value &&
                                          std::is_arithmetic<A>::value &&
                                          std::is_arithmetic<B>::value>::type>
inline typename std::common_type<A, B>::type max(A a, B b)
{
    typedef typename std::common_type<A, B>::type T;
    return ((T)a > (T)b) ? (T)a : (T)b;
}

template <class A, class B,
          class = typename std::enable_if<!std::is_same<A, B>::value &&
                                          std::is_arithmetic<A>::value &&
                                          std::is_arithmetic<B>::value>::type>
inline typename std::common_type<A, B>::type min(A a, B b)
{
    typedef typename std::common_type<A, B>::type T;
    return ((T)a < (T)b) ? (T)a : (T)b;
}
#endif
#endif
