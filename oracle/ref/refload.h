/* refload.h -- TEST INFRASTRUCTURE ONLY.
 *
 * In-process loader for the reference's prebuilt search engine
 * (/root/reference/www/bin/nrgrep_coords, an x86-64 non-PIE ELF shipped
 * without sources).  The executable's PT_LOAD segments are mapped at their
 * link addresses inside a PIE host process, its PLT/GOT is bound to the
 * host's libc, and its internal (unstripped) functions are then called
 * directly.  Nothing from the reference is copied into this repository: the
 * binary is read where it lies, at run time, in the build container only.
 *
 * Used by oracle/ref/difftest.c to pin the C restatement in oracle/ against
 * the reference function by function.
 */
#ifndef REFLOAD_H
#define REFLOAD_H
#include <stdint.h>

/* bind an import of the reference (e.g. "puts", "malloc") to fn instead of libc; call before ref_load */
void ref_override(const char *name, void *fn);
int ref_load(const char *path);          /* 0 on success */

/* addresses of unstripped symbols in nrgrep_coords (nm output) */
#define REF_searchPreproc   0x402570UL
#define REF_searchScan      0x402820UL
#define REF_searchFree      0x402770UL
#define REF_recPreproc      0x401ee0UL
#define REF_recFree         0x402000UL
#define REF_esimple_checkMatch1 0x414190UL

#define REF_OptDetWidth        (*(int *)0x621920UL)
#define REF_OptTransp          (*(int *)0x621924UL)
#define REF_OptSubs            (*(int *)0x621928UL)
#define REF_OptDel             (*(int *)0x62192cUL)
#define REF_OptIns             (*(int *)0x621930UL)
#define REF_OptRecSep          (*(char **)0x621938UL)
#define REF_OptRecPositive     (*(int *)0x621940UL)
#define REF_OptBufSize         (*(int *)0x621944UL)
#define REF_OptRecPos          (*(int *)0x621950UL)
#define REF_OptRecChar         (*(int *)0x621954UL)
#define REF_OptRecPatt         (*(char **)0x621958UL)
#define REF_RecByRec           (*(int *)0x6219b0UL)
#define REF_OptEndLine         (*(int *)0x6219c0UL)
#define REF_OptStartLine       (*(int *)0x6219c4UL)
#define REF_OptWholeRecord     (*(int *)0x6219c8UL)
#define REF_OptWholeWord       (*(int *)0x6219ccUL)
#define REF_OptErrors          (*(int *)0x6219d0UL)
#define REF_OptLiteral         (*(int *)0x6219d4UL)
#define REF_OptCaseInsensitive (*(int *)0x6219e8UL)

typedef void *(*ref_searchPreproc_t)(char *pattern);
typedef int (*ref_searchScan_t)(unsigned char **beg, unsigned char **end, void *P);
typedef void (*ref_void_t)(void);
typedef void (*ref_free_t)(void *);
#endif
