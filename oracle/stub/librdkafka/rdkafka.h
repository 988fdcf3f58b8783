/* Test-infrastructure stub — NOT product code.
 *
 * The reference driver header src/simulation/2DTissue.h:24 includes
 * "../utils/KafkaProducer.h", which includes <librdkafka/rdkafka.h>.  librdkafka is
 * not in this image and the configs under test say "no Kafka", so this header only
 * has to make KafkaProducer.h compile; none of these functions is ever reached
 * (kafkaEnabled == false).  No arithmetic of the reference is touched.
 */
#pragma once
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct rd_kafka_conf_s rd_kafka_conf_t;
typedef struct rd_kafka_s rd_kafka_t;
typedef struct rd_kafka_topic_s rd_kafka_topic_t;
typedef struct rd_kafka_topic_conf_s rd_kafka_topic_conf_t;
typedef enum { RD_KAFKA_PRODUCER, RD_KAFKA_CONSUMER } rd_kafka_type_t;
typedef enum { RD_KAFKA_CONF_UNKNOWN = -2, RD_KAFKA_CONF_INVALID = -1, RD_KAFKA_CONF_OK = 0 } rd_kafka_conf_res_t;
#define RD_KAFKA_PARTITION_UA ((int)-1)
#define RD_KAFKA_MSG_F_COPY 0x2
static inline rd_kafka_conf_t* rd_kafka_conf_new(void) { return NULL; }
static inline rd_kafka_conf_res_t rd_kafka_conf_set(rd_kafka_conf_t* c, const char* n, const char* v, char* e, size_t es)
{ (void)c; (void)n; (void)v; (void)e; (void)es; return RD_KAFKA_CONF_OK; }
static inline rd_kafka_t* rd_kafka_new(rd_kafka_type_t t, rd_kafka_conf_t* c, char* e, size_t es)
{ (void)t; (void)c; (void)e; (void)es; return NULL; }
static inline rd_kafka_topic_t* rd_kafka_topic_new(rd_kafka_t* rk, const char* t, rd_kafka_topic_conf_t* c)
{ (void)rk; (void)t; (void)c; return NULL; }
static inline int rd_kafka_flush(rd_kafka_t* rk, int ms) { (void)rk; (void)ms; return 0; }
static inline void rd_kafka_topic_destroy(rd_kafka_topic_t* t) { (void)t; }
static inline void rd_kafka_destroy(rd_kafka_t* rk) { (void)rk; }
static inline int rd_kafka_produce(rd_kafka_topic_t* t, int p, int f, void* pl, size_t l, const void* k, size_t kl, void* o)
{ (void)t; (void)p; (void)f; (void)pl; (void)l; (void)k; (void)kl; (void)o; return 0; }
#ifdef __cplusplus
}
#endif
