/* mallorn_ingest.h -- C-ABI of libmallorn_ingest.so: host ingest of the reference's light-curve CSV files.
 *
 * Serves `load_lightcurves` (src/utils/data_loader.py:36-62: pd.read_csv of every
 * split_XX/{train,test}_full_lightcurves.csv + pd.concat) together with the factorisation every extractor
 * repeats (`groupby('object_id')`, `Filter == band`; statistical.py:155,170): the result is exactly the flat,
 * file-order columns `mlrn_store_create` (mallorn_b200.h) takes, plus the id table.  Host-only (no CUDA);
 * plain pointers and sizes; 0 on success, negative MLRN_INGEST_E* otherwise with text from
 * mlrn_ingest_last_error().  Numbers are converted exactly as pandas' default C-engine converter does (see
 * csrc/ingest.cpp), so the columns are bit-identical to the reference's DataFrame.
 */
#ifndef MALLORN_INGEST_H
#define MALLORN_INGEST_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLRN_INGEST_OK 0
#define MLRN_INGEST_EINVAL (-1)    /* bad argument / more than 2^31-1 rows */
#define MLRN_INGEST_EIO (-2)       /* file cannot be opened or mapped, out of memory */
#define MLRN_INGEST_EFORMAT (-3)   /* missing column in a header, ragged line, non-numeric field */

typedef struct mlrn_table mlrn_table;   /* parsed columns, owned by the library until mlrn_table_free */

const char* mlrn_ingest_last_error(void);
/* Parse `n_paths` CSV files (header line naming object_id, Time (MJD), Flux, Flux_err, Filter in any order,
 * other columns ignored) as one concatenated table.  n_threads <= 0: all hardware threads. */
int mlrn_ingest_csv(const char* const* paths, int n_paths, int n_threads, mlrn_table** out);
int64_t mlrn_table_rows(const mlrn_table* t);
int32_t mlrn_table_objects(const mlrn_table* t);        /* distinct non-missing object ids */
const int32_t* mlrn_table_obj(const mlrn_table* t);     /* [rows] dense id by first appearance; -1 = missing id */
const uint8_t* mlrn_table_band(const mlrn_table* t);    /* [rows] 0..5 = u g r i z y, 6 = any other Filter */
const double* mlrn_table_time(const mlrn_table* t);     /* [rows] */
const double* mlrn_table_flux(const mlrn_table* t);
const double* mlrn_table_err(const mlrn_table* t);
int64_t mlrn_table_id_bytes(const mlrn_table* t);       /* id strings, concatenated */
const int64_t* mlrn_table_id_offsets(const mlrn_table* t);   /* [objects + 1] */
const char* mlrn_table_id_chars(const mlrn_table* t);
void mlrn_table_free(mlrn_table* t);


/* Factorisation of a DataFrame's string columns without Python objects: the host half of every extractor's
 * `lightcurves.groupby('object_id')` (statistical.py:155, colors.py:365, ...) and `Filter == band` masks (statistical.py:170)
 * when the caller hands over a pandas frame whose strings are Arrow-backed (pandas >= 3 default).  A column is given as
 * its Arrow arrays ("chunks", in row order): chunk c covers logical elements array_offset[c] .. array_offset[c]+length[c]
 * of its buffers -- `offsets` (int32 for string, int64 for large_string: offset_bytes 4 / 8), `data` (UTF-8 bytes) and
 * `validity` (bitmap, LSB first; the array of pointers or single entries may be NULL = no nulls).
 * mlrn_factorize_utf8: codes[row] = dense code of the row's string numbered by first appearance over all rows
 * (== pandas.factorize(sort=False)), -1 for a null; first_row[k] = row of the first appearance of code k (capacity: rows);
 * *n_unique = number of codes.  Threads over row ranges with a serial merge of the ranges' distinct strings.
 * mlrn_band_codes_utf8: band[row] = 0..5 for "u" "g" "r" "i" "z" "y", 6 for anything else or a null. */
int mlrn_factorize_utf8(int n_chunks, const void* const* offsets, const uint8_t* const* data, const uint8_t* const* validity,
                        const int64_t* array_offset, const int64_t* length, int offset_bytes, int n_threads,
                        int32_t* codes, int64_t* first_row, int32_t* n_unique);
int mlrn_band_codes_utf8(int n_chunks, const void* const* offsets, const uint8_t* const* data, const uint8_t* const* validity,
                         const int64_t* array_offset, const int64_t* length, int offset_bytes, int n_threads, uint8_t* band);

#ifdef __cplusplus
}
#endif
#endif /* MALLORN_INGEST_H */
