/* TEST INFRASTRUCTURE ONLY — stand-in for the reference's vkh helper library
 * (vkh/include/vkh.h).  Only the 22 vkh_* entry points that
 * src/vkvg_context.c / src/vkvg_context_internal.c call are declared; all are
 * implemented as host-memory fakes in oracle/ref_shim.c.  See oracle/Makefile. */
#ifndef ORACLE_STUB_VKH_H
#define ORACLE_STUB_VKH_H
#include <vulkan/vulkan.h>
#include <stdlib.h>
#include <stdio.h>
#include <assert.h>
#include <stdbool.h>
#include <stdint.h>
#include <string.h>
#include <math.h>

typedef int VkhMemoryUsage;
enum { VKH_MEMORY_USAGE_UNKNOWN = 0, VKH_MEMORY_USAGE_GPU_ONLY, VKH_MEMORY_USAGE_CPU_ONLY, VKH_MEMORY_USAGE_CPU_TO_GPU,
       VKH_MEMORY_USAGE_GPU_TO_CPU };

typedef struct _vkh_device_t *VkhDevice;
typedef struct _vkh_image_t  *VkhImage;
typedef struct _vkh_queue_t  *VkhQueue;
typedef struct _vkh_phy_t    *VkhPhyInfo;

struct _vkh_image_t { VkImageLayout layout; VkImage image; };
struct _vkh_queue_t { VkhDevice dev; uint32_t familyIndex; VkQueue queue; };

typedef struct _vkh_buffer_t {
    VkhDevice    pDev;
    VkBuffer     buffer;
    VkDeviceSize size;
    void        *mapped; /* host memory standing in for the persistently mapped VBO/IBO/UBO */
} vkh_buffer_t;
typedef vkh_buffer_t *VkhBuffer;

#define VK_CHECK_RESULT(f) { VkResult res__ = (f); assert(res__ == VK_SUCCESS); (void)res__; }

void  vkh_buffer_init(VkhDevice dev, VkBufferUsageFlags usage, VkhMemoryUsage mem, VkDeviceSize size, vkh_buffer_t *buff, bool mapped);
void  vkh_buffer_reset(vkh_buffer_t *buff);
void  vkh_buffer_resize(vkh_buffer_t *buff, VkDeviceSize newSize, bool mapped);
void *vkh_buffer_get_mapped_pointer(vkh_buffer_t *buff);
void  vkh_buffer_flush(vkh_buffer_t *buff);

void          vkh_cmd_begin(VkCommandBuffer cmd, VkCommandBufferUsageFlags flags);
void          vkh_cmd_end(VkCommandBuffer cmd);
void          vkh_cmd_buffs_create(VkhDevice dev, VkCommandPool pool, VkCommandBufferLevel level, uint32_t count, VkCommandBuffer *cmds);
VkCommandPool vkh_cmd_pool_create(VkhDevice dev, uint32_t qFamIndex, VkCommandPoolCreateFlags flags);
void          vkh_cmd_label_start(VkCommandBuffer cmd, const char *name, const float color[4]);
void          vkh_cmd_label_end(VkCommandBuffer cmd);
void          vkh_cmd_submit_timelined(VkhQueue q, VkCommandBuffer *cmd, VkSemaphore s, uint64_t w, uint64_t sig);
void          vkh_cmd_submit_timelined2(VkhQueue q, VkCommandBuffer *cmd, VkSemaphore s[2], uint64_t w[2], uint64_t sig[2]);
VkResult      vkh_timeline_wait(VkhDevice dev, VkSemaphore s, uint64_t v);
void          vkh_device_set_object_name(VkhDevice dev, VkObjectType t, uint64_t h, const char *name);
VkFence       vkh_fence_create_signaled(VkhDevice dev);

void     vkh_image_set_layout(VkCommandBuffer cmd, VkhImage img, VkImageAspectFlags aspect, VkImageLayout oldL,
                              VkImageLayout newL, VkPipelineStageFlags src, VkPipelineStageFlags dst);
void     vkh_image_destroy(VkhImage img);
VkImage  vkh_image_get_vkimage(VkhImage img);
VkhImage vkh_image_ms_create(VkhDevice dev, VkFormat format, VkSampleCountFlags samples, uint32_t w, uint32_t h,
                             VkhMemoryUsage mem, VkImageUsageFlags usage);
void     vkh_image_create_sampler(VkhImage img, VkFilter mag, VkFilter min, VkSamplerMipmapMode mip, VkSamplerAddressMode addr);
VkDescriptorImageInfo vkh_image_get_descriptor(VkhImage img, VkImageLayout layout);
#endif
