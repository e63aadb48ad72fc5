/* icw_plugin.h -- the reference's own transcode entry points, served by the B200 library.
 *
 * libicw_plugin.so exports the four symbols Winamp/XMPlay resolve by name in the reference DLL
 * (reference src/transcode.c:40-118), with the same signatures, ownership and error conventions:
 *
 *   open     -> handle (0 on any failure: unknown extension, bad header, unsupported format,
 *               output above ~2 GiB, sample rate above 2 MHz); fills *size (output bytes), *bps
 *               (16 or 24), *nch (always 2), *srate
 *   getData  -> bytes written into dest (whole frames only), 0 = end of file OR read error
 *   setTime  -> 1/0, seek to a millisecond position (clamped to the file, reference
 *               src/xwave_reader.c:782-817)
 *   close    -> releases the file; the DSP state (frame counter, Hilbert delay lines, dither stream)
 *               survives into the next file, as in the reference (src/config.c:171,174 defaults)
 *
 * The handle is a process-wide singleton like the reference's &the.mc_transcode (one transcode at
 * a time).  File parsing (RIFF/WAVE fmt 14/16/18/40 bytes, PCM / IEEE float / extensible GUIDs;
 * CWAVE V1/V2) follows reference src/xwave_reader.c:243-585 and src/cwave.h:47-84; the virtual
 * silence tail (sec_align) and the fades follow src/xwave_reader.c:593-728,838-904.
 * Hosts ask for 4-64 KB at a time (src/transcode.c:94); this layer reads ahead in large blocks and
 * runs them through libicw_b200.so, so a GPU launch is not paid per host call.  Read-ahead is invisible:
 * a seek or an early close puts the DSP state (frame counter, Hilbert memory, dither stream, counters)
 * back to exactly the frames the host took, which is all the reference ever advances it by.
 *
 * Configuration: icwp_load_config() reads the reference's own config file format; icwp_configure()
 * takes the chain description (filter, summation, DSP list, render settings; the per-file fields
 * fmt / n_channels / sample_rate / n_samples / fades are overwritten at open) plus the three
 * reader options the reference keeps in its config.
 */
#ifndef ICW_PLUGIN_H
#define ICW_PLUGIN_H

#include <stdint.h>
#include "icw_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/transcode.c:40 (narrow-character build), :82, :104, :113 */
intptr_t winampGetExtendedRead_open(const char *filename, int *size, int *bps, int *nch, int *srate);
intptr_t winampGetExtendedRead_getData(intptr_t handle, char *dest, int len, int *killswitch);
int      winampGetExtendedRead_setTime(intptr_t handle, int decode_pos_ms);
void     winampGetExtendedRead_close(intptr_t handle);

/* reader options of the reference config: SEC_ALIGN (s), FADE_IN / FADE_OUT (ms), src/config.c:147-152 */
typedef struct icwp_options {
    unsigned sec_align, fade_in_ms, fade_out_ms;
    int      clr_nframe_trk, clr_hilb_trk;      /* CLR_NFRAME_PT / CLR_HILB_PT, src/config.c:171,174 */
    int      device;                            /* CUDA device ordinal */
    int64_t  readahead_frames;                  /* 0 = default (1 Mi frames) */
} icwp_options;

/* chain = NULL restores the reference defaults (src/config.c:118-207); returns ICW_OK or ICW_E_* */
int  icwp_configure(const icw_chain_spec *chain, const icwp_options *opt);
/* fresh plugin state, like calling winampGetInModule2() again (src/in_cwave.c:551-572) */
void icwp_reset(void);
/* clips / peaks / reject counters of the transcode context (src/adv_modulator.c:445-465) */
int  icwp_stats(icw_stats *out);

/* how the overlapped reader did (N1: the step before the path, reference src/xwave_reader.c:838-904): file reads run
 * on a reader thread into page-locked blocks while the previous block is on the GPU */
typedef struct icwp_iostats {
    double   read_s;            /* time spent inside pread (reader thread, or the caller for the first block / after a seek) */
    double   wait_s;            /* time getData waited for a prefetched block that was not ready yet */
    double   gpu_s;             /* time inside icw_session_process_host */
    uint64_t read_bytes, frames;
    uint64_t blocks_prefetched, blocks_sync;
    uint64_t resettles;         /* seeks / early closes that had to put the DSP state back to what the host took */
} icwp_iostats;
int  icwp_io_stats(icwp_iostats *out, int reset);

/* header parsing alone (no GPU touched): what xwave_reader_create (src/xwave_reader.c:593-728)
 * would accept and report.  Returns 1 if the file is playable, 0 otherwise. */
typedef struct icwp_fileinfo {
    int      fmt;               /* ICW_FMT_* */
    int      n_channels;
    unsigned sample_rate;
    int64_t  n_samples;         /* frames in the file */
    int64_t  n_tail;            /* virtual silence frames appended for sec_align */
    int64_t  offset_data;       /* byte offset of the first frame */
    int64_t  n_fade_in, n_fade_out;
} icwp_fileinfo;
int  icwp_probe(const char *filename, const icwp_options *opt, icwp_fileinfo *out);

/* CWAVE data check (reference check_cwave, src/gui_cwave.c:82-130): CRC-32 of the sample data on the GPU.
 * Returns 1 when the data could be read; *has_crc = 0 for V1 files (no CRC in the header). */
int  icwp_check_cwave(const char *filename, uint32_t *crc_calc, uint32_t *crc_file, int *has_crc);

/* The reference's configuration file (text, "KEYWORD=values"; src/config.c:815-975, DSP-list lines
 * :562-774) -> chain + options, with the reference's acceptance rules: bounds clamped; an unknown
 * keyword, a line without '=' or version != 10 rejects the whole file (defaults left, returns 0) while a
 * value that does not parse is skipped, as in the reference; the DSP list is
 * taken only with one master at its head and after the "lock" fix-ups of amod_init
 * (src/adv_modulator.c:216-331).  Returns 1 when the file was accepted.  opt->device and
 * opt->readahead_frames are not config items and are preserved. */
int  icwp_load_config(const char *path, icw_chain_spec *chain, icwp_options *opt);
/* writes a file the reference's load_config() accepts (doubles as bit patterns, like save_config) */
int  icwp_save_config(const char *path, const icw_chain_spec *chain, const icwp_options *opt);

/* ---- the input-module table (reference Winamp/IN2.H:56-154, returned by winampGetInModule2, src/in_cwave.c:551-572) ----
 * Same member order and meaning as the SDK's In_Module for the narrow-character build (IN_VER 0x101); window and
 * module handles, the output module and the service pointer are opaque here.  What this library serves of it:
 * the description and extension list, Init / Quit (plugin-wide state like the reference's one-time init),
 * GetFileInfo (title = file name, length from the header incl. the sec_align tail, src/playback.c), IsOurFile,
 * GetLength.  The PLAYBACK half (Play / Pause / Stop / seek, the decode thread, visualisation feeds, src/playback.c)
 * is out of scope: those members are present and callable, Play returns ICWP_PLAY_UNSUPPORTED (a non-zero,
 * non -1 value: "other error" in the SDK's convention) and the rest do nothing. */
#define ICWP_IN_VER 0x101
#define ICWP_PLAY_UNSUPPORTED 2
typedef struct icwp_in_module {
    int version;
    char *description;
    void *hMainWindow, *hDllInstance;
    char *FileExtensions;
    int is_seekable;
    int UsesOutputPlug;
    void (*Config)(void *hwndParent);
    void (*About)(void *hwndParent);
    int  (*Init)(void);
    void (*Quit)(void);
    void (*GetFileInfo)(const char *file, char *title, int *length_in_ms);
    int  (*InfoBox)(const char *file, void *hwndParent);
    int  (*IsOurFile)(const char *fn);
    int  (*Play)(const char *fn);
    void (*Pause)(void);
    void (*UnPause)(void);
    int  (*IsPaused)(void);
    void (*Stop)(void);
    int  (*GetLength)(void);
    int  (*GetOutputTime)(void);
    void (*SetOutputTime)(int time_in_ms);
    void (*SetVolume)(int volume);
    void (*SetPan)(int pan);
    void (*SAVSAInit)(int maxlatency_in_ms, int srate);
    void (*SAVSADeInit)(void);
    void (*SAAddPCMData)(void *PCMData, int nch, int bps, int timestamp);
    int  (*SAGetMode)(void);
    int  (*SAAdd)(void *data, int timestamp, int csa);
    void (*VSAAddPCMData)(void *PCMData, int nch, int bps, int timestamp);
    int  (*VSAGetMode)(int *specNch, int *waveNch);
    int  (*VSAAdd)(void *data, int timestamp);
    void (*VSASetInfo)(int srate, int nch);
    int  (*dsp_isactive)(void);
    int  (*dsp_dosamples)(short int *samples, int numsamples, int bps, int nch, int srate);
    void (*EQSet)(int on, char data[10], int preamp);
    void (*SetInfo)(int bitrate, int srate, int stereo, int synched);
    void *outMod;
    void *service;
} icwp_in_module;
/* reference src/in_cwave.c:551: one-time plugin init (defaults, fresh contexts) + the table */
icwp_in_module *winampGetInModule2(void);

#ifdef __cplusplus
}
#endif
#endif
