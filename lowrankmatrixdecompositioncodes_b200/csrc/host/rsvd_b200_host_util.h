/* internal helpers shared by the two host translation units */
#ifndef RSVD_B200_HOST_UTIL_H
#define RSVD_B200_HOST_UTIL_H
#include <stddef.h>
void rsvd_api_error(const char *fmt, ...);     /* record a host-side error */
void rsvd_api_sync_error(void);                /* pull a pending device-layer error into the API status */
void rsvd_api_begin(void);                     /* clear status at the start of a TOP-LEVEL API call (no-op while nested) */
void rsvd_api_enter(void);                     /* begin + mark "inside a composite call": nested begins keep the status */
void rsvd_api_leave(void);
double *rsvd_upload(const double *h, size_t n);
void rsvd_download(double *h, const double *d, size_t n);
double *rsvd_host_calloc(size_t n);
double *rsvd_host_alloc_uninit(size_t n);      /* like rsvd_host_calloc but a recycled pinned block is not cleared */
void rsvd_host_free(double *p);
#endif
