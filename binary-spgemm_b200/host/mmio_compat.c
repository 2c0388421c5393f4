/* host/mmio_compat.c — see mmio_compat.h.  Behaviour mirrored from final/mmio.c:96-217 (what counts as a
 * valid banner, which error number each failure maps to, comment skipping before the size line). */
#include "mmio_compat.h"
#include <ctype.h>
#include <string.h>
#include <stdlib.h>

static void lower(char *s) { for (; *s; ++s) *s = (char)tolower((unsigned char)*s); }

struct kw { const char *word; char code; };
static int lookup(const struct kw *tab, const char *w, char *out)
{
    for (; tab->word; ++tab) if (strcmp(tab->word, w) == 0) { *out = tab->code; return 1; }
    return 0;
}

static const struct kw k_format[] = {{"coordinate", 'C'}, {"array", 'A'}, {0, 0}};
static const struct kw k_field[]  = {{"real", 'R'}, {"complex", 'C'}, {"pattern", 'P'}, {"integer", 'I'}, {0, 0}};
static const struct kw k_symm[]   = {{"general", 'G'}, {"symmetric", 'S'}, {"hermitian", 'H'}, {"skew-symmetric", 'K'}, {0, 0}};

int mm_read_banner(FILE *f, MM_typecode *matcode)
{
    char line[MM_MAX_LINE_LENGTH];
    char tok[5][MM_MAX_TOKEN_LENGTH];
    mm_clear_typecode(matcode);
    if (!fgets(line, sizeof line, f)) return MM_PREMATURE_EOF;
    if (sscanf(line, "%63s %63s %63s %63s %63s", tok[0], tok[1], tok[2], tok[3], tok[4]) != 5) return MM_PREMATURE_EOF;
    for (int i = 1; i < 5; ++i) lower(tok[i]);
    if (strncmp(tok[0], MatrixMarketBanner, strlen(MatrixMarketBanner)) != 0) return MM_NO_HEADER;
    if (strcmp(tok[1], "matrix") != 0) return MM_UNSUPPORTED_TYPE;
    (*matcode)[0] = 'M';
    if (!lookup(k_format, tok[2], &(*matcode)[1])) return MM_UNSUPPORTED_TYPE;
    if (!lookup(k_field,  tok[3], &(*matcode)[2])) return MM_UNSUPPORTED_TYPE;
    if (!lookup(k_symm,   tok[4], &(*matcode)[3])) return MM_UNSUPPORTED_TYPE;
    return 0;
}

int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz)
{
    char line[MM_MAX_LINE_LENGTH];
    *M = *N = *nz = 0;
    for (;;) {                                  /* skip '%' comment lines, then blank lines, until "M N nz" */
        if (!fgets(line, sizeof line, f)) return MM_PREMATURE_EOF;
        if (line[0] == '%') continue;
        if (sscanf(line, "%d %d %d", M, N, nz) == 3) return 0;
    }
}

int mm_is_valid(MM_typecode t)
{
    if (!mm_is_matrix(t)) return 0;
    if (mm_is_dense(t) && mm_is_pattern(t)) return 0;
    if (mm_is_real(t) && mm_is_hermitian(t)) return 0;
    if (mm_is_pattern(t) && (mm_is_hermitian(t) || mm_is_skew(t))) return 0;
    return 1;
}

static const char *rev(const struct kw *tab, char code)
{
    for (; tab->word; ++tab) if (tab->code == code) return tab->word;
    return NULL;
}

char *mm_typecode_to_str(MM_typecode t)
{
    const char *a = rev(k_format, t[1]), *b = rev(k_field, t[2]), *c = rev(k_symm, t[3]);
    if (!mm_is_matrix(t) || !a || !b || !c) return NULL;
    char *s = (char *)malloc(MM_MAX_LINE_LENGTH);
    if (s) snprintf(s, MM_MAX_LINE_LENGTH, "matrix %s %s %s", a, b, c);
    return s;
}

int mm_write_banner(FILE *f, MM_typecode t)
{
    char *s = mm_typecode_to_str(t);
    int ok = s && fprintf(f, "%s %s\n", MatrixMarketBanner, s) > 0;
    free(s);
    return ok ? 0 : MM_COULD_NOT_WRITE_FILE;
}

int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz)
{
    return fprintf(f, "%d %d %d\n", M, N, nz) > 0 ? 0 : MM_COULD_NOT_WRITE_FILE;
}
