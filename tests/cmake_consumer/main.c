/* Unmodified-consumer stand-in: plain C99 against <libsais.h> / <libsais64.h>. */
#include <libsais.h>
#include <libsais64.h>
#include <stdio.h>
#include <string.h>

int main(void)
{
    const char * text = "banana";
    int32_t SA[6];
    int32_t rc = libsais((const uint8_t *)text, SA, 6, 0, NULL);
    printf("libsais rc=%d", (int)rc);
    for (int i = 0; rc == 0 && i < 6; ++i) printf(" %d", (int)SA[i]);
    printf("\n");
    /* -2 (no usable GPU) is a legitimate answer of the library on a CPU-only box: there is no CPU fallback */
    return (rc == 0 && SA[0] == 5) || rc == -2 ? 0 : 1;
}
