/* oracle/ref_shim/lib/libimagequant.h -- TEST INFRASTRUCTURE.
 * libimagequant is not installed in this image and is only used by the reference's DECODE direction
 * (tileconv/colorquant.h:25, tileconv/colors.cpp:61-99), which is out of scope.  The reference's
 * colorquant.h only needs these three names to be declared for colors.cpp to compile. */
#ifndef TCB200_REF_SHIM_LIBIMAGEQUANT_H
#define TCB200_REF_SHIM_LIBIMAGEQUANT_H
typedef struct liq_attr liq_attr;
typedef struct liq_image liq_image;
typedef struct liq_result liq_result;
#endif
