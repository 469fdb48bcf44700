/* explore_extended.c -- TEST INFRASTRUCTURE ONLY (see oracle/NOTES_extended.md).
 *
 * Calls the reference's own searchPreproc @402570 in process (refload.c) for each pattern read from stdin and prints
 * what extendedPreproc @413260 decided: which scan routine was installed (simpleScan @416600 over a fixed-length
 * sub-pattern, or extendedScan @4116f0), where the verification is anchored and the window length.
 * Build: gcc -O1 -fPIE -pie -Ioracle/ref oracle/ref/explore_extended.c oracle/ref/refload.c -o oracle/_build/explore_extended -ldl
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "refload.h"
static int my_puts(const char *s) { (void)s; return 0; }
static void *zmalloc(size_t n) { return calloc(1, n ? n : 1); }
int main(int argc, char **argv)
{
    ref_override("puts", (void *)my_puts);
    ref_override("malloc", (void *)zmalloc);
    if (ref_load("/root/reference/www/bin/nrgrep_coords")) return 2;
    ((ref_void_t)REF_recPreproc)();
    REF_OptCaseInsensitive = 1; REF_OptErrors = 0; REF_OptIns = 1; REF_OptDel = 1; REF_OptSubs = 1; REF_OptTransp = 0;
    REF_OptStartLine = 0; REF_OptEndLine = 0;
    char line[1024];
    while (fgets(line, sizeof line, stdin)) {
        line[strcspn(line, "\n")] = 0;
        if (!line[0]) continue;
        char pat[1024]; strcpy(pat, line);
        long *sd = ((ref_searchPreproc_t)REF_searchPreproc)(pat);
        if (!sd) { printf("%s : preproc failed\n", line); continue; }
        int type = (int)sd[0];
        unsigned char *E = (unsigned char *)sd[1];
        printf("%-28s type=%d", line, type);
        if (type == 2 || type == 3) {   /* guess: extended */
            unsigned long scanfn = *(unsigned long *)E;
            int anchor = *(int *)(E + 0x2060), rest = *(int *)(E + 0x2064), vt = *(int *)(E + 0x2068);
            unsigned char *S = *(unsigned char **)(E + 0x10);
            printf(" scan=%lx anchor=%d rest=%d vtype=%d", scanfn, anchor, rest, vt);
            if (scanfn == 0x416600UL) printf(" simple: len=%d x=%d", *(int *)(S + 0x800), *(int *)(S + 0x804));
            else printf(" ext: wlen=%d m=%d", *(int *)(S + 0x1018), *(int *)(S + 0x101c));
        }
        printf("\n");
        ((ref_free_t)REF_searchFree)(sd);
    }
    return 0;
}
