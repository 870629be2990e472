/* Minimal libzmq 4.x ABI declaration, written for this repo.
 *
 * The image has libzmq (bundled with pyzmq) but no zmq.h; the reference's
 * bam2bam.c and the batched worker (integration/bwa_gpu_batch.c) are the only
 * files that include it.  This header declares just the functions and constants
 * they use so that the UNMODIFIED reference can be compiled into
 * integration/_host/ (see integration/Makefile).  Values are libzmq's public
 * ABI (stable since 3.2).  oracle/zmq_shim holds the checkers' own copy.
 */
#ifndef BWA_HOST_ZMQ_ABI_H
#define BWA_HOST_ZMQ_ABI_H
#include <stddef.h>
#include <errno.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ZMQ_HAUSNUMERO 156384712
#ifndef ETERM
#define ETERM (ZMQ_HAUSNUMERO + 53)
#endif

#define ZMQ_PAIR 0
#define ZMQ_PUB 1
#define ZMQ_SUB 2
#define ZMQ_REQ 3
#define ZMQ_REP 4
#define ZMQ_DEALER 5
#define ZMQ_ROUTER 6
#define ZMQ_PULL 7
#define ZMQ_PUSH 8

#define ZMQ_SUBSCRIBE 6
#define ZMQ_LINGER 17
#define ZMQ_SNDHWM 23
#define ZMQ_RCVHWM 24

#define ZMQ_DONTWAIT 1
#define ZMQ_SNDMORE 2

#define ZMQ_POLLIN 1
#define ZMQ_POLLOUT 2
#define ZMQ_POLLERR 4

typedef struct zmq_msg_t { unsigned char _[64]; } __attribute__((aligned(sizeof(void *)))) zmq_msg_t;
typedef void(zmq_free_fn)(void *data, void *hint);

typedef struct zmq_pollitem_t {
	void *socket;
	int fd;
	short events;
	short revents;
} zmq_pollitem_t;

int zmq_errno(void);
const char *zmq_strerror(int errnum);
void *zmq_init(int io_threads);
int zmq_term(void *context);
void *zmq_socket(void *context, int type);
int zmq_close(void *s);
int zmq_setsockopt(void *s, int option, const void *optval, size_t optvallen);
int zmq_bind(void *s, const char *addr);
int zmq_connect(void *s, const char *addr);
int zmq_send(void *s, const void *buf, size_t len, int flags);
int zmq_recv(void *s, void *buf, size_t len, int flags);
int zmq_poll(zmq_pollitem_t *items, int nitems, long timeout);
int zmq_msg_init(zmq_msg_t *msg);
int zmq_msg_init_size(zmq_msg_t *msg, size_t size);
int zmq_msg_init_data(zmq_msg_t *msg, void *data, size_t size, zmq_free_fn *ffn, void *hint);
int zmq_msg_send(zmq_msg_t *msg, void *s, int flags);
int zmq_msg_recv(zmq_msg_t *msg, void *s, int flags);
int zmq_msg_close(zmq_msg_t *msg);
void *zmq_msg_data(zmq_msg_t *msg);
size_t zmq_msg_size(zmq_msg_t *msg);

#ifdef __cplusplus
}
#endif
#endif
