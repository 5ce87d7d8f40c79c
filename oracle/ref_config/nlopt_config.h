/* Hand-written configuration header for compiling the UNMODIFIED reference
 * sources (under $(REF), normally /root/reference) into oracle/_ref/ with
 * plain gcc -- see oracle/Makefile.  The reference generates this file with
 * cmake from nlopt_config.h.in; we do not run its build system, so the
 * feature macros it would probe are simply stated here for
 * x86-64 Linux / glibc / gcc.  This is test infrastructure, not product.
 */
#ifndef ORACLE_REF_NLOPT_CONFIG_H
#define ORACLE_REF_NLOPT_CONFIG_H

#define MAJOR_VERSION 2
#define MINOR_VERSION 11
#define BUGFIX_VERSION 0

#define HAVE_COPYSIGN 1
#define HAVE_FPCLASSIFY 1
#define HAVE_GETPID 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_ISINF 1
#define HAVE_ISNAN 1
#define HAVE_STDINT_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_TIME 1
#define HAVE_UINT32_T 1
#define HAVE_UNISTD_H 1
#define TIME_WITH_SYS_TIME 1

#define SIZEOF_UNSIGNED_INT 4
#define SIZEOF_UNSIGNED_LONG 8

#define THREADLOCAL __thread

#endif
