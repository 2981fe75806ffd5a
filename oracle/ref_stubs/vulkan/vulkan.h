/* TEST INFRASTRUCTURE ONLY — stand-in for <vulkan/vulkan.h>.
 *
 * The reference's tessellation sources (src/vkvg_context.c,
 * src/vkvg_context_internal.c, src/vkvg_pattern.c) only touch Vulkan through
 * type names, a few enum constants and 17 function pointers
 * (src/vkvg_device_internal.c:38-58).  This header supplies just enough of
 * those names so that the sources compile where they lie under
 * /root/reference; oracle/ref_shim.c supplies recording implementations.
 * Nothing here is Vulkan: every handle is an opaque pointer and every enum an
 * int.  Used only to build oracle/_ref/ (see oracle/Makefile).
 */
#ifndef ORACLE_STUB_VULKAN_H
#define ORACLE_STUB_VULKAN_H
#include <stdint.h>
#include <stddef.h>

#define VKAPI_PTR
#define VKAPI_CALL
#define VKAPI_ATTR
#define VK_TRUE 1
#define VK_FALSE 0
#define VK_WHOLE_SIZE (~0ULL)
#define VK_NULL_HANDLE 0

typedef uint32_t VkFlags;
typedef uint32_t VkBool32;
typedef uint64_t VkDeviceSize;

#define STUB_HANDLE(n) typedef struct n##_T *n;
STUB_HANDLE(VkInstance) STUB_HANDLE(VkPhysicalDevice) STUB_HANDLE(VkDevice) STUB_HANDLE(VkQueue)
STUB_HANDLE(VkCommandBuffer) STUB_HANDLE(VkCommandPool) STUB_HANDLE(VkFence) STUB_HANDLE(VkSemaphore)
STUB_HANDLE(VkBuffer) STUB_HANDLE(VkImage) STUB_HANDLE(VkImageView) STUB_HANDLE(VkSampler)
STUB_HANDLE(VkDeviceMemory) STUB_HANDLE(VkFramebuffer) STUB_HANDLE(VkRenderPass) STUB_HANDLE(VkPipeline)
STUB_HANDLE(VkPipelineCache) STUB_HANDLE(VkPipelineLayout) STUB_HANDLE(VkDescriptorSet)
STUB_HANDLE(VkDescriptorSetLayout) STUB_HANDLE(VkDescriptorPool) STUB_HANDLE(VkSurfaceKHR)

typedef VkFlags VkSampleCountFlags;
typedef VkFlags VkImageAspectFlags;
typedef VkFlags VkBufferUsageFlags;
typedef VkFlags VkImageUsageFlags;
typedef VkFlags VkPipelineStageFlags;
typedef VkFlags VkShaderStageFlags;
typedef VkFlags VkStencilFaceFlags;
typedef VkFlags VkCommandBufferUsageFlags;
typedef VkFlags VkCommandPoolCreateFlags;
typedef VkFlags VkCommandBufferResetFlags;
typedef VkFlags VkFormatFeatureFlags;
typedef VkFlags VkDescriptorPoolCreateFlags;

typedef int VkResult;
#define VK_SUCCESS 0
#define VK_TIMEOUT 2

typedef int VkFormat;
#define VK_FORMAT_UNDEFINED 0
#define VK_FORMAT_R8G8B8A8_UNORM 37
#define VK_FORMAT_B8G8R8A8_UNORM 44
#define VK_FORMAT_S8_UINT 127

typedef int VkImageTiling;
#define VK_IMAGE_TILING_OPTIMAL 0
#define VK_IMAGE_TILING_LINEAR 1
typedef int VkImageLayout;
enum {
    VK_IMAGE_LAYOUT_UNDEFINED = 0, VK_IMAGE_LAYOUT_GENERAL, VK_IMAGE_LAYOUT_COLOR_ATTACHMENT_OPTIMAL,
    VK_IMAGE_LAYOUT_DEPTH_STENCIL_ATTACHMENT_OPTIMAL, VK_IMAGE_LAYOUT_DEPTH_STENCIL_READ_ONLY_OPTIMAL,
    VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL,
    VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL
};
typedef int VkFilter;
#define VK_FILTER_NEAREST 0
#define VK_FILTER_LINEAR 1
typedef int VkSamplerMipmapMode;
#define VK_SAMPLER_MIPMAP_MODE_NEAREST 0
typedef int VkSamplerAddressMode;
enum { VK_SAMPLER_ADDRESS_MODE_REPEAT = 0, VK_SAMPLER_ADDRESS_MODE_MIRRORED_REPEAT,
       VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_EDGE, VK_SAMPLER_ADDRESS_MODE_CLAMP_TO_BORDER };
typedef int VkAttachmentLoadOp;
typedef int VkPhysicalDeviceType;
typedef int VkPipelineBindPoint;
#define VK_PIPELINE_BIND_POINT_GRAPHICS 0
typedef int VkIndexType;
#define VK_INDEX_TYPE_UINT16 0
#define VK_INDEX_TYPE_UINT32 1
typedef int VkSubpassContents;
#define VK_SUBPASS_CONTENTS_INLINE 0
typedef int VkCommandBufferLevel;
#define VK_COMMAND_BUFFER_LEVEL_PRIMARY 0
typedef int VkStructureType;
enum { VK_STRUCTURE_TYPE_RENDER_PASS_BEGIN_INFO = 43, VK_STRUCTURE_TYPE_WRITE_DESCRIPTOR_SET = 35,
       VK_STRUCTURE_TYPE_DESCRIPTOR_SET_ALLOCATE_INFO = 34, VK_STRUCTURE_TYPE_DESCRIPTOR_POOL_CREATE_INFO = 33 };
typedef int VkDescriptorType;
#define VK_DESCRIPTOR_TYPE_COMBINED_IMAGE_SAMPLER 1
#define VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER 6
typedef int VkObjectType;
enum { VK_OBJECT_TYPE_FENCE = 7, VK_OBJECT_TYPE_BUFFER = 9, VK_OBJECT_TYPE_COMMAND_BUFFER = 6,
       VK_OBJECT_TYPE_DESCRIPTOR_SET = 23, VK_OBJECT_TYPE_DESCRIPTOR_POOL = 22, VK_OBJECT_TYPE_COMMAND_POOL = 25 };
typedef int VkDebugReportObjectTypeEXT;
#define VK_DEBUG_REPORT_OBJECT_TYPE_COMMAND_BUFFER_EXT 6

#define VK_SAMPLE_COUNT_1_BIT 1
#define VK_SAMPLE_COUNT_2_BIT 2
#define VK_SAMPLE_COUNT_4_BIT 4
#define VK_SAMPLE_COUNT_8_BIT 8
#define VK_IMAGE_ASPECT_COLOR_BIT 1
#define VK_IMAGE_ASPECT_DEPTH_BIT 2
#define VK_IMAGE_ASPECT_STENCIL_BIT 4
#define VK_STENCIL_FRONT_AND_BACK 3
#define VK_SHADER_STAGE_VERTEX_BIT 1
#define VK_COMMAND_BUFFER_USAGE_ONE_TIME_SUBMIT_BIT 1
#define VK_COMMAND_POOL_CREATE_RESET_COMMAND_BUFFER_BIT 2
#define VK_DESCRIPTOR_POOL_CREATE_FREE_DESCRIPTOR_SET_BIT 1
#define VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT 0x10
#define VK_BUFFER_USAGE_INDEX_BUFFER_BIT 0x40
#define VK_BUFFER_USAGE_VERTEX_BUFFER_BIT 0x80
#define VK_IMAGE_USAGE_TRANSFER_SRC_BIT 1
#define VK_IMAGE_USAGE_TRANSFER_DST_BIT 2
#define VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT 0x1
#define VK_PIPELINE_STAGE_FRAGMENT_SHADER_BIT 0x80
#define VK_PIPELINE_STAGE_EARLY_FRAGMENT_TESTS_BIT 0x100
#define VK_PIPELINE_STAGE_LATE_FRAGMENT_TESTS_BIT 0x200
#define VK_PIPELINE_STAGE_COLOR_ATTACHMENT_OUTPUT_BIT 0x400
#define VK_PIPELINE_STAGE_TRANSFER_BIT 0x1000
#define VK_PIPELINE_STAGE_BOTTOM_OF_PIPE_BIT 0x2000
#define VK_PIPELINE_STAGE_ALL_GRAPHICS_BIT 0x8000
#define VK_FORMAT_FEATURE_SAMPLED_IMAGE_BIT 0x1
#define VK_FORMAT_FEATURE_COLOR_ATTACHMENT_BIT 0x80
#define VK_FORMAT_FEATURE_BLIT_SRC_BIT 0x400
#define VK_FORMAT_FEATURE_BLIT_DST_BIT 0x800
#define VK_FORMAT_FEATURE_TRANSFER_SRC_BIT 0x4000
#define VK_FORMAT_FEATURE_TRANSFER_DST_BIT 0x8000

typedef struct { int32_t x, y; } VkOffset2D;
typedef struct { uint32_t width, height; } VkExtent2D;
typedef struct { int32_t x, y, z; } VkOffset3D;
typedef struct { uint32_t width, height, depth; } VkExtent3D;
typedef struct { VkOffset2D offset; VkExtent2D extent; } VkRect2D;
typedef struct { float x, y, width, height, minDepth, maxDepth; } VkViewport;
typedef union { float float32[4]; int32_t int32[4]; uint32_t uint32[4]; } VkClearColorValue;
typedef struct { float depth; uint32_t stencil; } VkClearDepthStencilValue;
typedef union { VkClearColorValue color; VkClearDepthStencilValue depthStencil; } VkClearValue;
typedef struct { VkImageAspectFlags aspectMask; uint32_t colorAttachment; VkClearValue clearValue; } VkClearAttachment;
typedef struct { VkRect2D rect; uint32_t baseArrayLayer; uint32_t layerCount; } VkClearRect;
typedef struct {
    VkStructureType sType; const void *pNext; VkRenderPass renderPass; VkFramebuffer framebuffer;
    VkRect2D renderArea; uint32_t clearValueCount; const VkClearValue *pClearValues;
} VkRenderPassBeginInfo;
typedef struct { VkImageAspectFlags aspectMask; uint32_t mipLevel, baseArrayLayer, layerCount; } VkImageSubresourceLayers;
typedef struct {
    VkImageSubresourceLayers srcSubresource; VkOffset3D srcOffset;
    VkImageSubresourceLayers dstSubresource; VkOffset3D dstOffset; VkExtent3D extent;
} VkImageCopy;
typedef VkImageCopy VkImageResolve;
typedef struct { VkSampler sampler; VkImageView imageView; VkImageLayout imageLayout; } VkDescriptorImageInfo;
typedef struct { VkBuffer buffer; VkDeviceSize offset, range; } VkDescriptorBufferInfo;
typedef struct {
    VkStructureType sType; const void *pNext; VkDescriptorSet dstSet; uint32_t dstBinding, dstArrayElement,
        descriptorCount; VkDescriptorType descriptorType; const VkDescriptorImageInfo *pImageInfo;
    const VkDescriptorBufferInfo *pBufferInfo; const void *pTexelBufferView;
} VkWriteDescriptorSet;
typedef struct { VkDescriptorType type; uint32_t descriptorCount; } VkDescriptorPoolSize;
typedef struct {
    VkStructureType sType; const void *pNext; VkDescriptorPoolCreateFlags flags; uint32_t maxSets, poolSizeCount;
    const VkDescriptorPoolSize *pPoolSizes;
} VkDescriptorPoolCreateInfo;
typedef struct {
    VkStructureType sType; const void *pNext; VkDescriptorPool descriptorPool; uint32_t descriptorSetCount;
    const VkDescriptorSetLayout *pSetLayouts;
} VkDescriptorSetAllocateInfo;
typedef struct { int unused; } VkPhysicalDeviceMemoryProperties;
typedef struct { int unused; } VkPhysicalDeviceFeatures;
typedef struct { int unused; } VkPhysicalDeviceVulkan12Features;
typedef struct { VkDeviceSize size; VkBufferUsageFlags usage; } VkBufferCreateInfo;
typedef struct { int unused; } VkAllocationCallbacks;
typedef struct { int unused; } VkCopyDescriptorSet;

/* function-pointer types for the 17 globals of src/vkvg_device_internal.c:38-58 */
typedef void (*PFN_vkCmdBindPipeline)(VkCommandBuffer, VkPipelineBindPoint, VkPipeline);
typedef void (*PFN_vkCmdBindDescriptorSets)(VkCommandBuffer, VkPipelineBindPoint, VkPipelineLayout, uint32_t, uint32_t,
                                            const VkDescriptorSet *, uint32_t, const uint32_t *);
typedef void (*PFN_vkCmdBindIndexBuffer)(VkCommandBuffer, VkBuffer, VkDeviceSize, VkIndexType);
typedef void (*PFN_vkCmdBindVertexBuffers)(VkCommandBuffer, uint32_t, uint32_t, const VkBuffer *, const VkDeviceSize *);
typedef void (*PFN_vkCmdDrawIndexed)(VkCommandBuffer, uint32_t, uint32_t, uint32_t, int32_t, uint32_t);
typedef void (*PFN_vkCmdDraw)(VkCommandBuffer, uint32_t, uint32_t, uint32_t, uint32_t);
typedef void (*PFN_vkCmdSetStencilCompareMask)(VkCommandBuffer, VkStencilFaceFlags, uint32_t);
typedef void (*PFN_vkCmdSetStencilReference)(VkCommandBuffer, VkStencilFaceFlags, uint32_t);
typedef void (*PFN_vkCmdSetStencilWriteMask)(VkCommandBuffer, VkStencilFaceFlags, uint32_t);
typedef void (*PFN_vkCmdBeginRenderPass)(VkCommandBuffer, const VkRenderPassBeginInfo *, VkSubpassContents);
typedef void (*PFN_vkCmdEndRenderPass)(VkCommandBuffer);
typedef void (*PFN_vkCmdSetViewport)(VkCommandBuffer, uint32_t, uint32_t, const VkViewport *);
typedef void (*PFN_vkCmdSetScissor)(VkCommandBuffer, uint32_t, uint32_t, const VkRect2D *);
typedef void (*PFN_vkCmdPushConstants)(VkCommandBuffer, VkPipelineLayout, VkShaderStageFlags, uint32_t, uint32_t,
                                       const void *);
typedef VkResult (*PFN_vkWaitForFences)(VkDevice, uint32_t, const VkFence *, VkBool32, uint64_t);
typedef VkResult (*PFN_vkResetFences)(VkDevice, uint32_t, const VkFence *);
typedef VkResult (*PFN_vkResetCommandBuffer)(VkCommandBuffer, VkCommandBufferResetFlags);

/* directly-called vk* entry points (implemented in oracle/ref_shim.c) */
void     vkCmdClearAttachments(VkCommandBuffer, uint32_t, const VkClearAttachment *, uint32_t, const VkClearRect *);
void     vkCmdCopyImage(VkCommandBuffer, VkImage, VkImageLayout, VkImage, VkImageLayout, uint32_t, const VkImageCopy *);
void     vkUpdateDescriptorSets(VkDevice, uint32_t, const VkWriteDescriptorSet *, uint32_t, const VkCopyDescriptorSet *);
VkResult vkCreateDescriptorPool(VkDevice, const VkDescriptorPoolCreateInfo *, const VkAllocationCallbacks *,
                                VkDescriptorPool *);
VkResult vkAllocateDescriptorSets(VkDevice, const VkDescriptorSetAllocateInfo *, VkDescriptorSet *);
VkResult vkFreeDescriptorSets(VkDevice, VkDescriptorPool, uint32_t, const VkDescriptorSet *);
void     vkDestroyDescriptorPool(VkDevice, VkDescriptorPool, const VkAllocationCallbacks *);
void     vkDestroyFence(VkDevice, VkFence, const VkAllocationCallbacks *);
void     vkFreeCommandBuffers(VkDevice, VkCommandPool, uint32_t, const VkCommandBuffer *);
void     vkDestroyCommandPool(VkDevice, VkCommandPool, const VkAllocationCallbacks *);
VkResult vkEndCommandBuffer(VkCommandBuffer);

#endif
