/* refload.c -- TEST INFRASTRUCTURE ONLY; see refload.h. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <elf.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <getopt.h>
#include "refload.h"

static void ref_trap(void) { fprintf(stderr, "refload: unresolved import called\n"); abort(); }

#define REF_MAX_OVERRIDES 16
static struct { const char *name; void *fn; } overrides[REF_MAX_OVERRIDES];
static int n_overrides;

void ref_override(const char *name, void *fn)
{
    if (n_overrides < REF_MAX_OVERRIDES) { overrides[n_overrides].name = name; overrides[n_overrides].fn = fn; n_overrides++; }
}
static void *ref_resolve(const char *name)
{
    for (int i = 0; i < n_overrides; i++)
        if (!strcmp(overrides[i].name, name)) return overrides[i].fn;
    return dlsym(RTLD_DEFAULT, name);
}

int ref_load(const char *path)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) { perror(path); return -1; }
    struct stat st;
    fstat(fd, &st);
    unsigned char *img = mmap(NULL, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (img == MAP_FAILED) { perror("mmap"); return -1; }
    Elf64_Ehdr *eh = (Elf64_Ehdr *)img;
    if (memcmp(eh->e_ident, ELFMAG, SELFMAG) || eh->e_type != ET_EXEC || eh->e_machine != EM_X86_64) {
        fprintf(stderr, "refload: %s is not an x86-64 ET_EXEC\n", path);
        return -1;
    }
    Elf64_Phdr *ph = (Elf64_Phdr *)(img + eh->e_phoff);
    Elf64_Dyn *dyn = NULL;
    for (int i = 0; i < eh->e_phnum; i++) {
        if (ph[i].p_type == PT_DYNAMIC) dyn = (Elf64_Dyn *)ph[i].p_vaddr;
        if (ph[i].p_type != PT_LOAD) continue;
        uintptr_t lo = ph[i].p_vaddr & ~0xfffUL;
        uintptr_t hi = (ph[i].p_vaddr + ph[i].p_memsz + 0xfff) & ~0xfffUL;
        void *m = mmap((void *)lo, hi - lo, PROT_READ | PROT_WRITE | PROT_EXEC,
                       MAP_PRIVATE | MAP_FIXED_NOREPLACE | MAP_ANONYMOUS, -1, 0);
        if (m != (void *)lo) { perror("refload: fixed mmap"); return -1; }
        memcpy((void *)ph[i].p_vaddr, img + ph[i].p_offset, ph[i].p_filesz);
    }
    if (!dyn) return -1;
    Elf64_Sym *symtab = NULL; char *strtab = NULL;
    Elf64_Rela *rela = NULL, *jmprel = NULL; size_t relasz = 0, pltrelsz = 0;
    for (Elf64_Dyn *d = dyn; d->d_tag != DT_NULL; d++) {
        switch (d->d_tag) {
        case DT_SYMTAB: symtab = (Elf64_Sym *)d->d_un.d_ptr; break;
        case DT_STRTAB: strtab = (char *)d->d_un.d_ptr; break;
        case DT_RELA: rela = (Elf64_Rela *)d->d_un.d_ptr; break;
        case DT_RELASZ: relasz = d->d_un.d_val; break;
        case DT_JMPREL: jmprel = (Elf64_Rela *)d->d_un.d_ptr; break;
        case DT_PLTRELSZ: pltrelsz = d->d_un.d_val; break;
        }
    }
    for (int pass = 0; pass < 2; pass++) {
        Elf64_Rela *r = pass ? jmprel : rela;
        size_t n = (pass ? pltrelsz : relasz) / sizeof(Elf64_Rela);
        for (size_t i = 0; i < n; i++) {
            const char *name = strtab + symtab[ELF64_R_SYM(r[i].r_info)].st_name;
            void *sym = ref_resolve(name);
            switch (ELF64_R_TYPE(r[i].r_info)) {
            case R_X86_64_JUMP_SLOT:
                *(void **)r[i].r_offset = sym ? sym : (void *)ref_trap;
                break;
            case R_X86_64_GLOB_DAT:
                *(void **)r[i].r_offset = sym;
                break;
            case R_X86_64_COPY:
                if (sym) memcpy((void *)r[i].r_offset, sym, symtab[ELF64_R_SYM(r[i].r_info)].st_size);
                break;
            }
        }
    }
    munmap(img, st.st_size);
    close(fd);
    return 0;
}
