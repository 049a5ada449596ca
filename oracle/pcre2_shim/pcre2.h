/* oracle/pcre2_shim/pcre2.h -- TEST INFRASTRUCTURE.
 *
 * The reference's sentence splitter (slimt/Regex.cc, slimt/Splitter.cc) includes <pcre2.h>.  This image carries
 * PCRE2's run-time library (libpcre2-8.so.0, 10.42) but not its development header, so the reference text front half
 * is compiled against this hand-written declaration of the dozen entry points and constants it uses (from PCRE2's
 * published API: pcre2api(3)) and linked against the run-time library itself.  Only oracle/Makefile uses it. */
#ifndef SLIMT_B200_ORACLE_PCRE2_SHIM_H
#define SLIMT_B200_ORACLE_PCRE2_SHIM_H
#include <stddef.h>
#include <stdint.h>

#if !defined(PCRE2_CODE_UNIT_WIDTH) || PCRE2_CODE_UNIT_WIDTH != 8
#error "the shim declares the 8-bit library only"
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef uint8_t PCRE2_UCHAR;
typedef const PCRE2_UCHAR *PCRE2_SPTR;
typedef size_t PCRE2_SIZE;
#define PCRE2_ZERO_TERMINATED (~(PCRE2_SIZE)0)

#define PCRE2_ANCHORED 0x80000000u
#define PCRE2_NO_UTF_CHECK 0x40000000u
#define PCRE2_DOTALL 0x00000020u
#define PCRE2_UTF 0x00080000u
#define PCRE2_NEWLINE_ANY 4
#define PCRE2_JIT_COMPLETE 0x00000001u
#define PCRE2_CONFIG_JIT 1

struct pcre2_real_code_8;
struct pcre2_real_match_data_8;
struct pcre2_real_compile_context_8;
struct pcre2_real_general_context_8;
struct pcre2_real_match_context_8;
typedef struct pcre2_real_code_8 pcre2_code;
typedef struct pcre2_real_match_data_8 pcre2_match_data;
typedef struct pcre2_real_compile_context_8 pcre2_compile_context;
typedef struct pcre2_real_general_context_8 pcre2_general_context;
typedef struct pcre2_real_match_context_8 pcre2_match_context;

pcre2_code *pcre2_compile_8(PCRE2_SPTR, PCRE2_SIZE, uint32_t, int *, PCRE2_SIZE *, pcre2_compile_context *);
void pcre2_code_free_8(pcre2_code *);
int pcre2_config_8(uint32_t, void *);
int pcre2_jit_compile_8(pcre2_code *, uint32_t);
int pcre2_get_error_message_8(int, PCRE2_UCHAR *, PCRE2_SIZE);
pcre2_match_data *pcre2_match_data_create_from_pattern_8(const pcre2_code *, pcre2_general_context *);
void pcre2_match_data_free_8(pcre2_match_data *);
int pcre2_match_8(const pcre2_code *, PCRE2_SPTR, PCRE2_SIZE, PCRE2_SIZE, uint32_t, pcre2_match_data *, pcre2_match_context *);
PCRE2_SIZE *pcre2_get_ovector_pointer_8(pcre2_match_data *);
PCRE2_SIZE pcre2_get_startchar_8(pcre2_match_data *);

#define pcre2_compile pcre2_compile_8
#define pcre2_code_free pcre2_code_free_8
#define pcre2_config pcre2_config_8
#define pcre2_jit_compile pcre2_jit_compile_8
#define pcre2_get_error_message pcre2_get_error_message_8
#define pcre2_match_data_create_from_pattern pcre2_match_data_create_from_pattern_8
#define pcre2_match_data_free pcre2_match_data_free_8
#define pcre2_match pcre2_match_8
#define pcre2_get_ovector_pointer pcre2_get_ovector_pointer_8
#define pcre2_get_startchar pcre2_get_startchar_8

#ifdef __cplusplus
}
#endif
#endif
