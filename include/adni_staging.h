/*
 * adni_staging — C-ABI of the host-side input staging for the B200 training path.
 *
 * The reference feeds the path from `MultiModalDataset.__getitem__` (pkg/utils/dataloader.py:197-321): 32 DataLoader
 * worker processes each `nib.load(path).get_fdata()` a gzip-compressed NIfTI-1 volume into a float64 numpy array,
 * normalise it on the CPU (two sorts per scan) and pickle it back to the trainer.  Here the host side only DECODES:
 * worker threads of this library inflate the .nii(.gz) files straight into caller-owned (pinned) buffers as fp32
 * intensities / uint8 brain masks; one cudaMemcpyAsync later the normalisation runs on the GPU
 * (adni_quantile_minmax_normalize / adni_standardize in adni_b200.h).
 *
 * nibabel (the reference's reader; nibabel==4.0.2, environment.yml) is a third-party dependency absent from the
 * reference tree: these entry points restate its documented behaviour for single-file NIfTI-1 images -
 *   nib.load(p).get_fdata()  ==  raw.astype(float64) * scl_slope + scl_inter   (no scaling when scl_slope is 0 or
 *   not finite), array axes (dim[1], dim[2], dim[3]) with dim[1] fastest ON DISK;
 *   torch.tensor(get_fdata()) is C-contiguous, i.e. dim[3] fastest IN MEMORY -
 * so every reader below writes element (i, j, k) to dst[(i*dim2 + j)*dim3 + k].
 *
 * Conventions: host pointers only, no CUDA dependency (this is a separate shared library, libadni_stage.so);
 * returns 0 or a negative ADNI_E* code (same values as adni_b200.h); adni_stage_last_error() explains; thread-safe
 * (no global mutable state except the thread-local error string).
 */
#ifndef ADNI_STAGING_H
#define ADNI_STAGING_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADNI_STAGE_OK 0
#define ADNI_STAGE_EINVAL (-1)  /* bad argument / buffer too small                                      */
#define ADNI_STAGE_ENOTSUP (-2) /* not a single-file NIfTI-1 image, unsupported datatype, non-binary mask */
#define ADNI_STAGE_EIO (-5)     /* cannot open / truncated / corrupt gzip stream (reference: nibabel ImageFileError) */

typedef struct {
  int32_t ndim;      /* dim[0] with trailing singleton axes dropped (3 for the reference's volumes)      */
  int64_t dim[7];    /* dim[1..7]                                                                        */
  int32_t datatype;  /* NIfTI-1 datatype code (2 u8, 4 i16, 8 i32, 16 f32, 64 f64, 256 i8, 512 u16, 768 u32, 1024 i64, 1280 u64) */
  int32_t bitpix;
  int32_t swapped;   /* 1 = file is in the other byte order                                              */
  int32_t scaled;    /* 1 = get_fdata applies scl_slope / scl_inter                                      */
  double scl_slope, scl_inter;
  int64_t vox_offset;
  int64_t nvox;      /* product of dim[0..ndim)                                                          */
} adni_nifti_info;

const char* adni_stage_last_error(void);
int adni_stage_version(void);

/* Header only (reads 352 bytes).  Replaces nib.load(path).header / .shape (dataloader.py:206, 226, 240). */
int adni_nifti_read_info(const char* path, adni_nifti_info* info);

/* dst[nvox] = float64 get_fdata(), C order.  Bit-exact restatement of the reference's read (parity tests). */
int adni_nifti_read_f64(const char* path, double* dst, int64_t capacity, adni_nifti_info* info);

/* dst[nvox] = (float) get_fdata(), C order: the staging format of the GPU normalisation kernels.  Exact whenever
 * the stored datatype is u8/i8/i16/u16/f32 without scaling (every image of the reference's ANTs/MNI pipeline);
 * otherwise the fp32 rounding of the fp64 value.  Replaces dataloader.py:206-207 (PET) and :226-227 (MRI). */
int adni_nifti_read_f32(const char* path, float* dst, int64_t capacity, adni_nifti_info* info);

/* dst[nvox] = 1 where get_fdata() != 0 else 0, C order.  The reference multiplies by the mask image
 * (dataloader.py:239-247); a mask with values other than 0 and 1 would scale intensities there and cannot be
 * represented here: ADNI_STAGE_ENOTSUP. */
int adni_nifti_read_mask_u8(const char* path, uint8_t* dst, int64_t capacity, adni_nifti_info* info);

/* Decode n files with `threads` worker threads: file i goes to (char*)dst + i * stride_bytes; kind 0 = fp32
 * volume, 1 = uint8 mask.  Every file must have exactly `expect_nvox` voxels (the batch is one tensor).  status[i]
 * receives the per-file code; returns the first non-zero one.  A NULL path leaves its slot untouched (absent
 * modality).  Replaces the DataLoader worker pool of train_anat_cnn.py:187-198. */
int adni_stage_volumes(const char* const* paths, int n, int kind, void* dst, int64_t stride_bytes, int64_t expect_nvox,
                       int threads, int* status);

#ifdef __cplusplus
}
#endif
#endif /* ADNI_STAGING_H */
